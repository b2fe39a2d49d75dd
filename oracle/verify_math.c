/* oracle/verify_math.c -- TEST INFRASTRUCTURE.  Exhaustive accuracy sweep of
 * marlnav_trig.h against double-precision libm over EVERY float32 in the domain
 * (sin/cos: [-pi, pi]; acos: [-1, 1]).  Prints max error in ulps of the exact
 * result and the fraction of inputs that are not correctly rounded.
 *   gcc -O2 -mfma -ffp-contract=off -fopenmp verify_math.c -lm -o _verify_math */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "marlnav_trig.h"
#include "torch_cpu_math.h"

static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline double ulp_of(double exact) {
    float f = (float)fabs(exact);
    if (f == 0.0f) return ldexp(1.0, -149);
    int e; frexpf(f, &e);
    return ldexp(1.0, e - 24 < -149 ? -149 : e - 24);
}
typedef struct { double max_ulp; uint64_t n, not_cr; float worst; } acc_t;
static void upd(acc_t* a, float x, float got, double exact) {
    double err = fabs((double)got - exact) / ulp_of(exact);
    a->n++;
    if (got != (float)exact) a->not_cr++;
    if (err > a->max_ulp) { a->max_ulp = err; a->worst = x; }
}
static void merge(acc_t* d, const acc_t* s) {
    d->n += s->n; d->not_cr += s->not_cr;
    if (s->max_ulp > d->max_ulp) { d->max_ulp = s->max_ulp; d->worst = s->worst; }
}
int main(void) {
    const uint32_t pi_bits = 0x40490fdbu, one_bits = 0x3f800000u;
    acc_t S = {0}, C = {0}, A = {0}, SS = {0}, SC = {0}, SA = {0};
#pragma omp parallel
    {
        acc_t s = {0}, c = {0}, a = {0}, ss = {0}, sc = {0}, sa = {0};
#pragma omp for schedule(static) nowait
        for (int64_t i = 0; i <= (int64_t)pi_bits; ++i)
            for (int sg = 0; sg < 2; ++sg) {
                float x = u2f((uint32_t)i | (sg ? 0x80000000u : 0));
                float sn, cs; mt_sincosf(x, &sn, &cs);
                upd(&s, x, sn, sin((double)x)); upd(&c, x, cs, cos((double)x));
                if ((i & 15) == 0) { upd(&ss, x, tcm_sinf(x), sin((double)x)); upd(&sc, x, tcm_cosf(x), cos((double)x)); }
            }
#pragma omp for schedule(static) nowait
        for (int64_t i = 0; i <= (int64_t)one_bits; ++i)
            for (int sg = 0; sg < 2; ++sg) {
                float x = u2f((uint32_t)i | (sg ? 0x80000000u : 0));
                upd(&a, x, mt_acosf(x), acos((double)x));
                if ((i & 15) == 0) upd(&sa, x, tcm_acosf(x), acos((double)x));
            }
#pragma omp critical
        { merge(&S, &s); merge(&C, &c); merge(&A, &a); merge(&SS, &ss); merge(&SC, &sc); merge(&SA, &sa); }
    }
    printf("marlnav sin : n=%llu max_ulp=%.4f (x=%.9g) not_correctly_rounded=%.4f%%\n", (unsigned long long)S.n, S.max_ulp, S.worst, 100.0 * S.not_cr / S.n);
    printf("marlnav cos : n=%llu max_ulp=%.4f (x=%.9g) not_correctly_rounded=%.4f%%\n", (unsigned long long)C.n, C.max_ulp, C.worst, 100.0 * C.not_cr / C.n);
    printf("marlnav acos: n=%llu max_ulp=%.4f (x=%.9g) not_correctly_rounded=%.4f%%\n", (unsigned long long)A.n, A.max_ulp, A.worst, 100.0 * A.not_cr / A.n);
    printf("sleef-u10 sin (1/16 sample): max_ulp=%.4f not_cr=%.4f%%\n", SS.max_ulp, 100.0 * SS.not_cr / SS.n);
    printf("sleef-u10 cos (1/16 sample): max_ulp=%.4f not_cr=%.4f%%\n", SC.max_ulp, 100.0 * SC.not_cr / SC.n);
    printf("sleef-u10 acos(1/16 sample): max_ulp=%.4f not_cr=%.4f%%\n", SA.max_ulp, 100.0 * SA.not_cr / SA.n);
    return 0;
}
