// scripts/ubench/pair2_math.cuh -- MEASURED AND REJECTED (round 2); kept as evidence with pair2.cu.
//
// Two (agent, object) pairs per lane in one instruction stream: the arithmetic of
// marlnav_b200/csrc/marlnav_math.cuh and pair_finish on sm_100a's packed float32x2 operations
// (add/sub/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2, operands in aligned register pairs;
// a scalar register broadcasts to both halves as an operand mode, negation is an operand
// modifier).  The idea: the step kernels are bound by instruction issue, every add / mul / fma
// of the pair evaluation (/root/reference/marlnav/environment.py:271-286) exists once per pair,
// so two pairs side by side would cost one issue slot per arithmetic step instead of two.
//
// Result on B200 (scripts/ubench/pair2.cu, 303 104 threads x 24 objects x 40 repetitions):
// bit-identical to the scalar sequences on all 14.5 M values, 26 % fewer warp instructions
// (417.6 M vs 562.7 M) -- and only 8 % less time (546 vs 592 us).  ncu: issue slots 69 % busy,
// FMA-pipe instructions = 34 % of the cycles, and 69 + 34 = 103: a packed instruction holds the
// issue port for TWO cycles (stall reasons on the FFMA2 lines: math-pipe throttle 45 %,
// not_selected 37 %).  FFMA2 keeps the FP32 pipe's flop rate (scripts/ubench/f32x2.cu: 73.8 vs
// 71.9 TFLOP/s) but frees no issue slots, so it cannot help an issue-bound kernel; the 8 % came
// from the scalar savings that were then moved into marlnav_math.cuh (FMNMX3, clamp-based zero
// fix).  Two facts worth keeping for anyone who tries again:
//   * ptxas contracts  mul.rn.f32x2 -> add.rn.f32x2  into FFMA2 even though both carry .rn and
//     the file is compiled -fmad=false (it never does that for scalar .rn ops).  A sum that must
//     stay unfused and has a product as an operand has to be written  fma(x, ONE, y)  with
//     ONE = {1.0f, 1.0f} read from kernel arguments at run time.
//   * negation has no packed PTX form; negating the halves with scalar neg.f32 is folded by
//     ptxas into the consumer's operand modifier (FFMA2 -R.F32x2...).
#pragma once
#include <cuda_runtime.h>

#include "marlnav_math.cuh"       // -I marlnav_b200/csrc

namespace mn {

typedef unsigned long long f32x2;       // {lo, hi} in one aligned 64-bit register pair

__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d;
}
__device__ __forceinline__ void unpk2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 dup2(float x) { return pk2(x, x); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ f32x2 neg2(f32x2 a) {
    float lo, hi; unpk2(a, lo, hi); return pk2(-lo, -hi);
}
// x + y with a product among the operands, kept unfused (see the header)
__device__ __forceinline__ f32x2 add2_unfused(f32x2 x, f32x2 y, f32x2 one_rt) { return fma2(x, one_rt, y); }

__device__ __forceinline__ float rsqrt_approx(float x) {
    float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
// sqrt_rn_nonzero on both halves (x in [2^-100, 2^100])
__device__ __forceinline__ f32x2 sqrt2_rn_nonzero(f32x2 x) {
    float xl, xh; unpk2(x, xl, xh);
    const f32x2 y = pk2(rsqrt_approx(xl), rsqrt_approx(xh));
    const f32x2 g = mul2(x, y), h = mul2(y, dup2(0.5f));
    return fma2(fma2(neg2(g), g, x), h, g);
}
// sqrt_rn_normal on both halves (x == 0 or x in [2^-100, 2^100]; MUFU input clamped from below)
__device__ __forceinline__ f32x2 sqrt2_rn_normal(f32x2 x) {
    float xl, xh; unpk2(x, xl, xh);
    const f32x2 y = pk2(rsqrt_approx(fmaxf(xl, 7.888609052210118e-31f)), rsqrt_approx(fmaxf(xh, 7.888609052210118e-31f)));
    const f32x2 g = mul2(x, y), h = mul2(y, dup2(0.5f));
    return fma2(fma2(neg2(g), g, x), h, g);
}
// div2_rn_normal on both halves: (a/b, c/b)
__device__ __forceinline__ void div22_rn_normal(f32x2 a, f32x2 c, f32x2 b, f32x2& qa, f32x2& qc) {
    float bl, bh; unpk2(b, bl, bh);
    f32x2 r = pk2(rcp_approx(bl), rcp_approx(bh));
    const f32x2 nb = neg2(b);
    r = fma2(r, fma2(nb, r, dup2(1.0f)), r);
    const f32x2 q0 = fma2(a, r, dup2(0.0f)), q1 = fma2(c, r, dup2(0.0f));
    qa = fma2(r, fma2(nb, q0, a), q0);
    qc = fma2(r, fma2(nb, q1, c), q1);
}
// rcp_rn_normal on both halves
__device__ __forceinline__ f32x2 rcp2_rn_normal(f32x2 b) {
    float bl, bh; unpk2(b, bl, bh);
    f32x2 r = pk2(rcp_approx(bl), rcp_approx(bh));
    const f32x2 nb = neg2(b);
    r = fma2(r, fma2(nb, r, dup2(1.0f)), r);
    return fma2(r, fma2(nb, r, dup2(1.0f)), r);
}

// acos_f on both halves (|x| <= 1)
__device__ __forceinline__ void acos2_f(float xl, float xh, float& rl, float& rh) {
    const f32x2 x = pk2(xl, xh);
    const bool sl = fabsf(xl) <= 0.5f, sh = fabsf(xh) <= 0.5f;
    float xxl, xxh, whl, whh;
    unpk2(mul2(x, x), xxl, xxh);
    unpk2(mul2(pk2(1.0f - fabsf(xl), 1.0f - fabsf(xh)), dup2(0.5f)), whl, whh);
    const f32x2 z = pk2(sl ? xxl : whl, sh ? xxh : whh);
    float ql, qh; unpk2(sqrt2_rn_normal(z), ql, qh);
    const f32x2 t = pk2(sl ? xl : ql, sh ? xh : qh);
    f32x2 u = fma2(dup2(+0.4197454825e-1f), z, dup2(+0.2424046025e-1f));
    u = fma2(u, z, dup2(+0.4547423869e-1f));
    u = fma2(u, z, dup2(+0.7495029271e-1f));
    u = fma2(u, z, dup2(+0.1666677296e+0f));
    const f32x2 as = fma2(mul2(t, z), u, t);
    float smalll, smallh, twl, twh, bigl, bigh;
    unpk2(sub2(dup2(MN_PIO2_HI), sub2(as, dup2(MN_PIO2_LO))), smalll, smallh);
    const f32x2 twice = add2(as, as);
    unpk2(twice, twl, twh);
    unpk2(sub2(dup2(MN_PI_HI), sub2(twice, dup2(MN_PI_LO))), bigl, bigh);
    rl = sl ? smalll : (xl < 0.0f ? bigl : twl);
    rh = sh ? smallh : (xh < 0.0f ? bigh : twh);
}

// geom_fast for two objects at once (ex, ey = object - own, packed over the two objects).
// `lo` accumulates min |component| as in the scalar version; `hisum` accumulates the SUM of the
// d^2 (all >= 0, so sum < T implies every d^2 < T; a NaN or inf poisons it) with one packed add.
__device__ __forceinline__ void geom2_fast(f32x2 ex, f32x2 ey, f32x2& d, f32x2& nx, f32x2& ny, float& lo, f32x2& hisum) {
    const f32x2 dd = fma2(ey, ey, mul2(ex, ex));
    float exl, exh, eyl, eyh;
    unpk2(ex, exl, exh); unpk2(ey, eyl, eyh);
    lo = min3_nan_abs(lo, exl, eyl);
    lo = min3_nan_abs(lo, exh, eyh);
    hisum = add2(hisum, dd);
    d = sqrt2_rn_nonzero(dd);
    div22_rn_normal(ex, ey, d, nx, ny);
}

// pair_finish for two pairs: headings (hx, hy) packed over the same two pairs
__device__ __forceinline__ void pair2_finish(f32x2 d, f32x2 nx, f32x2 ny, f32x2 hx, f32x2 hy, float cap, f32x2 one_rt,
                                             float& angl, float& angh) {
    float dl, dh, dotl, doth, nxl, nxh, ml, mh, acl, ach;
    unpk2(add2_unfused(mul2(hx, nx), mul2(hy, ny), one_rt), dotl, doth);
    dotl = clamp_nan(dotl, -1.0f, 1.0f); doth = clamp_nan(doth, -1.0f, 1.0f);
    acos2_f(dotl, doth, acl, ach);
    unpk2(mul2(pk2(dotl, doth), hx), ml, mh);
    unpk2(nx, nxl, nxh); unpk2(d, dl, dh);
    float al = nxl > ml ? -acl : acl, ah = nxh > mh ? -ach : ach;
    if (dl < cap) al = 0.0f;
    if (dh < cap) ah = 0.0f;
    angl = al; angh = ah;
}

}  // namespace mn
