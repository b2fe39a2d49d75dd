/*
 * oracle/torch_cpu_math.h -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * Bit-exact restatement of the three float32 transcendentals the reference's
 * environment step reaches through torch's CPU backend:
 *
 *   torch.cos / torch.sin   /root/reference/marlnav/environment.py:134-135
 *   torch.acos              /root/reference/marlnav/environment.py:286
 *
 * torch's vectorised CPU kernels (AVX2 / AVX512 builds, which is what
 * torch.backends.cpu.get_cpu_capability() reports on the hosts used here) call
 * the third-party library SLEEF, statically linked into libtorch_cpu.so
 * (symbols Sleef_sinf8_u10 / Sleef_cosf8_u10 / Sleef_acosf8_u10 and the 16-lane
 * variants).  SLEEF is NOT vendored in /root/reference (it is a dependency of
 * the dependency `torch==2.5.1`, requirements.txt:4; the container has torch
 * 2.11.0).  What follows restates SLEEF's published u10 ("1.0 ULP") algorithms
 * for the FMA-enabled targets: double-float ("df") arithmetic, Cody-Waite
 * reduction with the 3-part PI split, and the minimax polynomials.
 *
 * Pinning: tests/test_oracle_math.py compares these functions against
 * torch.cos/sin/acos bit for bit (sampled in the CPU suite; the exhaustive
 * sweep over every float32 in [-pi, pi] resp. [-1, 1] is
 * oracle/verify_math_exhaustive.py, result recorded in DESIGN.md).
 *
 * Domain: |x| < 125 for sin/cos (the step clamps turn angles to [-pi, pi]);
 * outside that SLEEF switches to Payne-Hanek, which is not restated -- the
 * functions below return NaN there so misuse is loud.
 *
 * Compile with -ffp-contract=off so that only the explicit fmaf() calls fuse.
 */
#ifndef MARLNAV_ORACLE_TORCH_CPU_MATH_H
#define MARLNAV_ORACLE_TORCH_CPU_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct { float x, y; } tcm_f2;  /* double-float: value = x + y, |y| <= ulp(x)/2 */

static inline tcm_f2 tcm_mk(float x, float y) { tcm_f2 r; r.x = x; r.y = y; return r; }

static inline float tcm_mla(float a, float b, float c) { return fmaf(a, b, c); }

static inline float tcm_mulsign(float x, float y) {
    uint32_t ux, uy; memcpy(&ux, &x, 4); memcpy(&uy, &y, 4);
    ux ^= (uy & 0x80000000u);
    float r; memcpy(&r, &ux, 4); return r;
}

/* (x + y) for |x| >= |y| */
static inline tcm_f2 tcm_add_f_f(float x, float y) {
    float s = x + y;
    return tcm_mk(s, (x - s) + y);
}
/* (x + y), no ordering assumption */
static inline tcm_f2 tcm_add2_f_f(float x, float y) {
    float s = x + y;
    float v = s - x;
    return tcm_mk(s, (x - (s - v)) + (y - v));
}
static inline tcm_f2 tcm_add_f2_f(tcm_f2 x, float y) {
    float s = x.x + y;
    return tcm_mk(s, ((x.x - s) + y) + x.y);
}
static inline tcm_f2 tcm_add2_f2_f(tcm_f2 x, float y) {
    float s = x.x + y;
    float v = s - x.x;
    float t = (x.x - (s - v)) + (y - v);
    return tcm_mk(s, t + x.y);
}
static inline tcm_f2 tcm_add_f_f2(float x, tcm_f2 y) {
    float s = x + y.x;
    return tcm_mk(s, ((x - s) + y.x) + y.y);
}
static inline tcm_f2 tcm_add2_f_f2(float x, tcm_f2 y) {
    float s = x + y.x;
    float v = s - x;
    return tcm_mk(s, ((x - (s - v)) + (y.x - v)) + y.y);
}
static inline tcm_f2 tcm_sub_f2_f2(tcm_f2 x, tcm_f2 y) {
    float s = x.x - y.x;
    float t = x.x - s;
    t = t - y.x;
    t = t + x.y;
    return tcm_mk(s, t - y.y);
}
static inline tcm_f2 tcm_scale(tcm_f2 d, float s) { return tcm_mk(d.x * s, d.y * s); }

/* FMA flavours of the df products */
static inline tcm_f2 tcm_mul_f_f(float x, float y) {
    float t = x * y;
    return tcm_mk(t, fmaf(x, y, -t));
}
static inline tcm_f2 tcm_mul_f2_f2(tcm_f2 x, tcm_f2 y) {
    float t = x.x * y.x;
    return tcm_mk(t, fmaf(x.x, y.y, fmaf(x.y, y.x, fmaf(x.x, y.x, -t))));
}
static inline float tcm_mul_f2_f2_f(tcm_f2 x, tcm_f2 y) {
    return fmaf(x.x, y.x, fmaf(x.y, y.x, x.x * y.y));
}
static inline tcm_f2 tcm_squ(tcm_f2 x) {
    float t = x.x * x.x;
    return tcm_mk(t, fmaf(x.x + x.x, x.y, fmaf(x.x, x.x, -t)));
}
static inline tcm_f2 tcm_rec_f(float d) {
    float t = 1.0f / d;
    return tcm_mk(t, t * fmaf(-d, t, 1.0f));
}
static inline tcm_f2 tcm_sqrt_f(float d) {
    float t = sqrtf(d);
    return tcm_scale(tcm_mul_f2_f2(tcm_add2_f_f2(d, tcm_mul_f_f(t, t)), tcm_rec_f(t)), 0.5f);
}

#define TCM_PI_A2f 3.1414794921875f
#define TCM_PI_B2f 0.00011315941810607910156f
#define TCM_PI_C2f 1.9841872589410058936e-09f
#define TCM_1_PIf  0.318309886183790671537767526745028724f
#define TCM_TRIGRANGEMAX2f 125.0f

static inline float tcm_sincos_poly(tcm_f2 s_in) {
    tcm_f2 t = s_in;
    tcm_f2 s = tcm_squ(s_in);
    float u = 2.6083159809786593541503e-06f;
    u = tcm_mla(u, s.x, -0.0001981069071916863322258f);
    u = tcm_mla(u, s.x, 0.00833307858556509017944336f);
    tcm_f2 x = tcm_add_f_f2(1.0f,
        tcm_mul_f2_f2(tcm_add_f_f(-0.166666597127914428710938f, u * s.x), s));
    return tcm_mul_f2_f2_f(t, x);
}

static inline float tcm_sinf(float d) {
    if (!(fabsf(d) < TCM_TRIGRANGEMAX2f)) return NAN;
    float u = rintf(d * TCM_1_PIf);
    int q = (int)rintf(u);
    float v = tcm_mla(u, -TCM_PI_A2f, d);
    tcm_f2 s = tcm_add2_f_f(v, u * (-TCM_PI_B2f));
    s = tcm_add_f2_f(s, u * (-TCM_PI_C2f));
    float r = tcm_sincos_poly(s);
    if (q & 1) r = -r;
    if (d == 0.0f && signbit(d)) r = d;
    return r;
}

static inline float tcm_cosf(float d) {
    if (!(fabsf(d) < TCM_TRIGRANGEMAX2f)) return NAN;
    float dq = tcm_mla(rintf(tcm_mla(d, TCM_1_PIf, -0.5f)), 2.0f, 1.0f);
    int q = (int)rintf(dq);
    tcm_f2 s = tcm_add2_f_f(d, dq * (-TCM_PI_A2f * 0.5f));
    s = tcm_add2_f2_f(s, dq * (-TCM_PI_B2f * 0.5f));
    s = tcm_add2_f2_f(s, dq * (-TCM_PI_C2f * 0.5f));
    float r = tcm_sincos_poly(s);
    if ((q & 2) == 0) r = -r;
    return r;
}

static inline float tcm_acosf(float d) {
    float ad = fabsf(d);
    int o = ad < 0.5f;
    float x2 = o ? (d * d) : ((1.0f - ad) * 0.5f);
    tcm_f2 x = o ? tcm_mk(ad, 0.0f) : tcm_sqrt_f(x2);
    if (ad == 1.0f) x = tcm_mk(0.0f, 0.0f);

    float u = +0.4197454825e-1f;
    u = tcm_mla(u, x2, +0.2424046025e-1f);
    u = tcm_mla(u, x2, +0.4547423869e-1f);
    u = tcm_mla(u, x2, +0.7495029271e-1f);
    u = tcm_mla(u, x2, +0.1666677296e+0f);
    u = u * (x2 * x.x);

    tcm_f2 y = tcm_sub_f2_f2(tcm_mk(3.1415927410125732422f / 2, -8.7422776573475857731e-08f / 2),
                             tcm_add_f_f(tcm_mulsign(x.x, d), tcm_mulsign(u, d)));
    x = tcm_add_f2_f(x, u);
    if (!o) y = tcm_scale(x, 2.0f);
    if (!o && d < 0.0f)
        y = tcm_sub_f2_f2(tcm_mk(3.1415927410125732422f, -8.7422776573475857731e-08f), y);
    return y.x + y.y;
}

#endif
