/*
 * oracle/marlnav_trig.h -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * The float32 sin/cos/acos the oracle (and, operation for operation, the CUDA
 * kernel in marlnav_b200/csrc/marlnav_math.cuh) use for
 *   torch.cos / torch.sin   /root/reference/marlnav/environment.py:134-135
 *   torch.acos              /root/reference/marlnav/environment.py:286
 *
 * Why not "what torch does": on the MKL builds of torch used here the three ops
 * run in closed-source MKL VML (vsCos/vsSin/vsAcos, high-accuracy mode); there
 * is no published algorithm to restate and its bits differ from torch's other
 * CPU backend (SLEEF u10, restated bit-exactly in torch_cpu_math.h) on 2-9 % of
 * inputs by 1 ulp.  Any faithful implementation is therefore "as close to the
 * reference as the reference is to itself across builds".  The functions below
 * are built ONLY from IEEE-754 correctly-rounded operations (+ - * fma sqrt), so
 * C on the host and CUDA on the device produce identical bits, and they are
 * cheap on the device (the step is instruction-issue bound).
 *
 * Accuracy (oracle/verify_math.c, every float32 in the domain,
 * against the correctly rounded result): see DESIGN.md "Transcendentals".
 *
 *   mt_sincosf(t): |t| <= pi (the step clamps turn angles first).  Cody-Waite
 *       reduction by pi/2 (k in -2..2, two-term split, fused), cephes-style
 *       minimax polynomials on [-pi/4, pi/4], quadrant fix-up.
 *   mt_acosf(x):   |x| <= 1.  acos(|x|) = sqrt(1-|x|) * P7(|x|) (the form of Abramowitz &
 *       Stegun 4.4.46; coefficients refitted for relative error and tuned over float32
 *       neighbours), pi - that for x < 0.  (1-|x|) is exact for |x| >= 1/2, so small
 *       angles keep full relative accuracy.  Max error 2.10 ulp: 4 instructions per
 *       call cheaper on the device than the 1.12-ulp two-range asin form of round 1,
 *       which bought 5.5 % of the (8,16) step and 3.7 % of the (3,3) step.
 *
 * Compile with -ffp-contract=off: only the explicit fmaf() calls may fuse.
 */
#ifndef MARLNAV_ORACLE_TRIG_H
#define MARLNAV_ORACLE_TRIG_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#define MT_2_PI    0.636619746685028076171875f      /* RN(2/pi) */
#define MT_PIO2_HI 1.57079637050628662109375f       /* RN(pi/2) */
#define MT_PIO2_LO -4.37113882867379114031791687e-8f /* RN(pi/2 - MT_PIO2_HI) */
#define MT_PI_HI   3.1415927410125732421875f        /* RN(pi) */
#define MT_PI_LO   -8.74227765734758228063583374e-8f /* RN(pi - MT_PI_HI) */

static inline void mt_sincosf(float t, float* sn, float* cs) {
    const float k = rintf(t * MT_2_PI);
    float r = fmaf(k, -MT_PIO2_HI, t);
    r = fmaf(k, -MT_PIO2_LO, r);
    const float z = r * r;
    float ps = fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
    ps = fmaf(ps, z, -1.6666654611e-1f);
    const float s = (t == 0.0f) ? t : fmaf(r * z, ps, r);   /* keeps sin(-0) = -0 like torch */
    float pc = fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
    pc = fmaf(pc, z, 4.166664568298827e-2f);
    const float c = fmaf(z, fmaf(z, pc, -0.5f), 1.0f);
    const int q = (int)k & 3;                        /* two's complement: -1 -> 3, -2 -> 2 */
    const float s1 = (q & 1) ? c : s;
    const float c1 = (q & 1) ? s : c;
    *sn = (q & 2) ? -s1 : s1;
    *cs = ((q + 1) & 2) ? -c1 : c1;
}

#define MT_ACOS_C0 1.57079625f
#define MT_ACOS_C1 -0.214598596f
#define MT_ACOS_C2 0.0889772698f
#define MT_ACOS_C3 -0.0501640774f
#define MT_ACOS_C4 0.0308625922f
#define MT_ACOS_C5 -0.0170451012f
#define MT_ACOS_C6 0.00663866755f
#define MT_ACOS_C7 -0.00125347136f

static inline float mt_acosf(float x) {
    const float a = fabsf(x);
    const float t = sqrtf(1.0f - a);                 /* exact difference for a >= 1/2 */
    float p = MT_ACOS_C7;
    p = fmaf(p, a, MT_ACOS_C6);
    p = fmaf(p, a, MT_ACOS_C5);
    p = fmaf(p, a, MT_ACOS_C4);
    p = fmaf(p, a, MT_ACOS_C3);
    p = fmaf(p, a, MT_ACOS_C2);
    p = fmaf(p, a, MT_ACOS_C1);
    p = fmaf(p, a, MT_ACOS_C0);
    const float r = t * p;                           /* acos(|x|) */
    return x < 0.0f ? (MT_PI_HI - (r - MT_PI_LO)) : r;
}

/* ln(x) for x in [2^-24, 1] -- the Box-Muller radius of the noisy agent reset
 * (/root/reference/marlnav/utils.py:381-385 draws its position noise from
 * MultivariateNormal, i.e. torch.randn on the CPU generator; here the normals are
 * addressed Philox draws, SURVEY.md Appendix D, so they need a logarithm both sides
 * can reproduce).  Cephes-style logf: x = m 2^e with m in [sqrt(1/2), sqrt(2)),
 * f = m - 1, ln(1+f) = f - f^2/2 + f^3 P(f), e ln2 split in two.  IEEE-only
 * (integer exponent extraction, + - * fma).  Max error 0.83 ulp over all 2^24
 * inputs k 2^-24 the sampler can produce (oracle/verify_math.c). */
static inline float mt_logf01(float x) {
    uint32_t b; memcpy(&b, &x, 4);
    int e = (int)((b >> 23) & 0xffu) - 126;          /* x = m 2^e, m in [1/2, 1) */
    b = (b & 0x007fffffu) | 0x3f000000u;
    float m; memcpy(&m, &b, 4);
    float f;
    if (m < 0.707106781186547524f) { e -= 1; f = (m + m) - 1.0f; } else { f = m - 1.0f; }
    const float z = f * f;
    float p = 7.0376836292e-2f;
    p = fmaf(p, f, -1.1514610310e-1f);
    p = fmaf(p, f, 1.1676998740e-1f);
    p = fmaf(p, f, -1.2420140846e-1f);
    p = fmaf(p, f, 1.4249322787e-1f);
    p = fmaf(p, f, -1.6668057665e-1f);
    p = fmaf(p, f, 2.0000714765e-1f);
    p = fmaf(p, f, -2.4999993993e-1f);
    p = fmaf(p, f, 3.3333331174e-1f);
    const float fe = (float)e;
    float y = (f * z) * p;
    y = fmaf(fe, -2.12194440e-4f, y);
    y = fmaf(z, -0.5f, y);
    return fmaf(fe, 0.693359375f, f + y);
}

/* Two independent standard normals from two uniforms, u1 in (0, 1], u2 in [0, 1):
 * r = sqrt(-2 ln u1), theta = 2 pi (u2 - 1/2) in [-pi, pi). */
static inline void mt_box_muller(float u1, float u2, float* z0, float* z1) {
    const float rad = sqrtf(-2.0f * mt_logf01(u1));
    float sn, cs;
    mt_sincosf(6.2831854820251465f * (u2 - 0.5f), &sn, &cs);
    *z0 = rad * cs; *z1 = rad * sn;
}

#endif
