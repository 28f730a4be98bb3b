/*
 * ode_pow.h -- pow() and exp() that are BIT-IDENTICAL to the host's glibc 2.39 libm
 * (x86-64, the FMA ifunc variant every AVX2+FMA CPU selects), usable from C, C++ and CUDA.
 *
 * Why this exists: the reference's PI step-size controller computes err.powf(alpha) and
 * fac_old.powf(-beta) on every attempted step (src/controller.rs:53-54) and hinit() one more
 * (src/dopri5.rs:236, src/dop853.rs:244); f64::powf is libm `pow`. The result sets h_new, so
 * a device pow that is merely accurate (CUDA's is <= 2 ulp; glibc's is not correctly rounded
 * either) makes h differ in the last bit about once per 10^3 calls, which chaotic systems
 * (Lorenz) amplify until accept/reject decisions flip. Identical step counts for every
 * trajectory therefore need the same bits, i.e. the same algorithm, tables and fusions.
 *
 * Algorithm: the public table-driven pow/exp of ARM Optimized Routines (as adopted by glibc
 * >= 2.28): log(x) in double-double via a 128-entry table + degree-7 polynomial, y*log(x)
 * with an FMA-exact product, exp via a 128-entry 2^(i/128) table + degree-5 polynomial.
 * The tables are read from the running libm.so.6 (tools/extract_libm_tables.py ->
 * ode_pow_tables.h). Which multiply-adds are FUSED was taken from the machine code of
 * glibc's FMA build (gcc -mfma contracts a*b+c): every fma() below is one vfmadd there, every
 * separate * and + is a separate vmulsd/vaddsd. nvcc is run with -fmad=false so nothing
 * else fuses on the device; gcc hosts must use -ffp-contract=off.
 *
 * Upstream and licence: the algorithm is that of ARM Optimized Routines (math/pow.c, math/exp.c, math/powf.c;
 * MIT OR Apache-2.0 WITH LLVM-exception), which glibc ships under the LGPL-2.1-or-later in sysdeps/ieee754.
 * This file is a restatement written from the published algorithm and the observed machine code, not a copy of
 * either source; the numeric tables are data read at build time from the libm installed on the build host.
 *
 * tests/test_pow_clone.py (CPU) and tests/test_gpu_pow.py (device) check bit equality with
 * libm over the controller's domain, the full double range and all special cases.
 */
#ifndef ODE_POW_H
#define ODE_POW_H

#include "ode_pow_tables.h"

#if defined(__CUDACC__)
#define ODE_POW_FN __host__ __device__ static __forceinline__
#define ODE_POW_SLOW __host__ __device__ static __noinline__
#else
#include <math.h>
#include <stdint.h>
#include <string.h>
#define ODE_POW_FN static inline
#define ODE_POW_SLOW static __attribute__((noinline))
#endif

#if defined(__CUDA_ARCH__)
#define ODE_ASU64(d) ((unsigned long long)__double_as_longlong(d))
#define ODE_ASF64(u) (__longlong_as_double((long long)(u)))
#else
ODE_POW_FN unsigned long long ode_asu64_(double d) {
    unsigned long long u;
    memcpy(&u, &d, 8);
    return u;
}
ODE_POW_FN double ode_asf64_(unsigned long long u) {
    double d;
    memcpy(&d, &u, 8);
    return d;
}
#define ODE_ASU64(d) ode_asu64_(d)
#define ODE_ASF64(u) ode_asf64_(u)
#endif

typedef unsigned long long ode_u64;
typedef long long ode_i64;
typedef unsigned int ode_u32;

/* Table words: lt = 128 x {invc, logc, logctail}, et = 256 exp-table words, both as raw
 * 64-bit patterns. On the device they live in shared memory (per-lane data-dependent
 * indices would serialise in the constant cache); on the host they are static arrays. */
typedef struct ode_pow_tabs {
    const ode_u64* lt;
    const ode_u64* et;
} ode_pow_tabs;

#if !defined(__CUDA_ARCH__)
static const ode_u64 ODE_POW_LOG_TAB_HOST[ODE_POW_LOG_TAB_WORDS] = ODE_POW_LOG_TAB_INIT;
static const ode_u64 ODE_EXP_TAB_HOST[ODE_EXP_TAB_WORDS] = ODE_EXP_TAB_INIT;
ODE_POW_FN ode_pow_tabs ode_pow_host_tabs(void) {
    ode_pow_tabs t;
    t.lt = ODE_POW_LOG_TAB_HOST;
    t.et = ODE_EXP_TAB_HOST;
    return t;
}
#endif

/* Scalar constants of the two polynomials. On the device they are read from the constant
 * bank (one LDCU / c[] operand) instead of being materialised as 64-bit immediates (two
 * UMOV each), which measurably costs issue slots in the FP64-bound step loop. */
#if defined(__CUDACC__)
static __constant__ double c_ode_pow_k[19] = {
    ODE_POW_LN2HI, ODE_POW_LN2LO, ODE_POW_A0, ODE_POW_A1, ODE_POW_A2, ODE_POW_A3, ODE_POW_A4, ODE_POW_A5,
    ODE_POW_A6, ODE_EXP_INVLN2N, ODE_EXP_SHIFT, ODE_EXP_NEGLN2HIN, ODE_EXP_NEGLN2LON, ODE_EXP_C2, ODE_EXP_C3,
    ODE_EXP_C4, ODE_EXP_C5, 0x1p1009, 0x1p-1022};
#endif
#if defined(__CUDA_ARCH__)
#define ODE_K(i, lit) c_ode_pow_k[i]
#else
#define ODE_K(i, lit) (lit)
#endif
#define K_LN2HI ODE_K(0, ODE_POW_LN2HI)
#define K_LN2LO ODE_K(1, ODE_POW_LN2LO)
#define K_A0 ODE_K(2, ODE_POW_A0)
#define K_A1 ODE_K(3, ODE_POW_A1)
#define K_A2 ODE_K(4, ODE_POW_A2)
#define K_A3 ODE_K(5, ODE_POW_A3)
#define K_A4 ODE_K(6, ODE_POW_A4)
#define K_A5 ODE_K(7, ODE_POW_A5)
#define K_A6 ODE_K(8, ODE_POW_A6)
#define K_INVLN2N ODE_K(9, ODE_EXP_INVLN2N)
#define K_SHIFT ODE_K(10, ODE_EXP_SHIFT)
#define K_NEGLN2HIN ODE_K(11, ODE_EXP_NEGLN2HIN)
#define K_NEGLN2LON ODE_K(12, ODE_EXP_NEGLN2LON)
#define K_C2 ODE_K(13, ODE_EXP_C2)
#define K_C3 ODE_K(14, ODE_EXP_C3)
#define K_C4 ODE_K(15, ODE_EXP_C4)
#define K_C5 ODE_K(16, ODE_EXP_C5)

#define ODE_POW_OFF 0x3fe6955500000000ULL
#define ODE_POW_SIGN_BIAS 0x40000u /* 0x800 << 7 */

/* exp's handling of a scale whose exponent over/underflowed (|x| >= 512 ... 1024) */
ODE_POW_SLOW double ode_exp_specialcase_(double tmp, ode_u64 sbits, ode_u64 ki) {
    double scale, y;
    if ((ki & 0x80000000ULL) == 0) {
        sbits -= 1009ULL << 52;
        scale = ODE_ASF64(sbits);
        y = 0x1p1009 * fma(scale, tmp, scale);
        return y;
    }
    sbits += 1022ULL << 52;
    scale = ODE_ASF64(sbits);
    double st = scale * tmp; /* NOT fused in libm: the product is reused below */
    y = scale + st;
    if (fabs(y) < 1.0) {
        double hi, lo, one = 1.0;
        if (y < 0.0) one = -1.0;
        lo = scale - y + st;
        hi = one + y;
        lo = one - hi + y + lo;
        y = (hi + lo) - one;
        if (y == 0.0) y = ODE_ASF64(sbits & 0x8000000000000000ULL);
    }
    return 0x1p-1022 * y;
}

/* The main path of exp_inline below as straight-line code: exp(x + xtail) * (-1)^(sign_bias != 0)
 * for 2^-54 <= |x| < 512 (ode_exp_in_main_range_), where no special case applies and the scale
 * needs no adjustment. The step-loop kernels run it speculatively and test the range afterwards. */
ODE_POW_FN int ode_exp_in_main_range_(double x) {
    return ((ode_u32)(ODE_ASU64(x) >> 52) & 0x7ff) - 0x3c9u < 0x3fu;
}
ODE_POW_FN double ode_exp_main_(double x, double xtail, ode_u32 sign_bias, ode_pow_tabs T) {
    double kd = fma(x, K_INVLN2N, K_SHIFT);
    ode_u64 ki = ODE_ASU64(kd);
    kd -= K_SHIFT;
    double r = fma(kd, K_NEGLN2HIN, x);
    r = fma(kd, K_NEGLN2LON, r);
    r += xtail;
    ode_u32 idx = 2u * (ode_u32)(ki & 127u);
    ode_u64 top = (ki + sign_bias) << 45;
    double tail = ODE_ASF64(T.et[idx]);
    ode_u64 sbits = T.et[idx + 1] + top;
    double r2 = r * r;
    double p01 = fma(r, K_C3, K_C2);
    double p23 = fma(r, K_C5, K_C4);
    double tmp = fma(p01, r2, tail + r);
    tmp = fma(r2 * r2, p23, tmp);
    double scale = ODE_ASF64(sbits);
    return fma(scale, tmp, scale);
}

/* exp(x + xtail) * (-1)^(sign_bias != 0); shared by pow and exp. */
ODE_POW_FN double ode_exp_inline_(double x, double xtail, ode_u32 sign_bias, int has_tail, ode_pow_tabs T) {
    ode_u32 abstop = (ode_u32)(ODE_ASU64(x) >> 52) & 0x7ff;
    if (abstop - 0x3c9u >= 0x3fu) {
        if (abstop - 0x3c9u >= 0x80000000u) {
            double one = 1.0 + x; /* tiny |x| */
            return sign_bias ? -one : one;
        }
        if (abstop >= 0x409u) {
            if (!has_tail) { /* plain exp(): inf / nan / overflow */
                if (ODE_ASU64(x) == 0xfff0000000000000ULL) return 0.0;
                if (abstop >= 0x7ffu) return 1.0 + x;
            }
            double big = (ODE_ASU64(x) >> 63) ? 0x1p-767 : 0x1p769;
            double s = sign_bias ? -big : big;
            return s * big; /* __math_uflow / __math_oflow */
        }
        abstop = 0;
    }
    double kd = fma(x, K_INVLN2N, K_SHIFT);
    ode_u64 ki = ODE_ASU64(kd);
    kd -= K_SHIFT;
    double r = fma(kd, K_NEGLN2HIN, x);
    r = fma(kd, K_NEGLN2LON, r);
    if (has_tail) r += xtail;
    ode_u32 idx = 2u * (ode_u32)(ki & 127u);
    ode_u64 top = (ki + sign_bias) << 45;
    double tail = ODE_ASF64(T.et[idx]);
    ode_u64 sbits = T.et[idx + 1] + top;
    double r2 = r * r;
    double p01 = fma(r, K_C3, K_C2);
    double p23 = fma(r, K_C5, K_C4);
    double tmp = fma(p01, r2, tail + r);
    tmp = fma(r2 * r2, p23, tmp);
    if (abstop == 0) return ode_exp_specialcase_(tmp, sbits, ki);
    double scale = ODE_ASF64(sbits);
    return fma(scale, tmp, scale);
}

/* 0 = not an integer, 1 = odd integer, 2 = even integer (iy: bits of a finite non-zero y) */
ODE_POW_FN int ode_pow_checkint_(ode_u64 iy) {
    int e = (int)(iy >> 52) & 0x7ff;
    if (e < 0x3ff) return 0;
    if (e > 0x3ff + 52) return 2;
    if (iy & ((1ULL << (0x3ff + 52 - e)) - 1)) return 0;
    if (iy & (1ULL << (0x3ff + 52 - e))) return 1;
    return 2;
}

/*
 * Everything outside "x normal positive, 2^-65 <= |y| < 2^63". Returns 1 with *res set when
 * the result is final, 0 when the main path should continue with the adjusted *pix /
 * *psign_bias (negative x with integer y, subnormal x).
 */
ODE_POW_SLOW int ode_pow_special_(double x, double y, ode_u64* pix, ode_u32* psign_bias, double* res) {
    ode_u64 ix = *pix, iy = ODE_ASU64(y);
    ode_u32 topx = (ode_u32)(ix >> 52), topy = (ode_u32)(iy >> 52);
    ode_u32 sign_bias = 0;
    /* zeroinfnan(i) = 2*i - 1 >= 2*asuint64(INFINITY) - 1 */
    if (2 * iy - 1 >= 2 * 0x7ff0000000000000ULL - 1) {
        if (2 * iy == 0) { *res = 1.0; return 1; } /* (signalling NaN x: also quiet 1.0 here) */
        if (ix == 0x3ff0000000000000ULL) { *res = 1.0; return 1; }
        if (2 * ix > 2 * 0x7ff0000000000000ULL || 2 * iy > 2 * 0x7ff0000000000000ULL) { *res = x + y; return 1; }
        if (2 * ix == 2 * 0x3ff0000000000000ULL) { *res = 1.0; return 1; }
        if ((2 * ix < 2 * 0x3ff0000000000000ULL) == !(iy >> 63)) { *res = 0.0; return 1; }
        *res = y * y;
        return 1;
    }
    if (2 * ix - 1 >= 2 * 0x7ff0000000000000ULL - 1) {
        double x2 = x * x;
        if ((ix >> 63) && ode_pow_checkint_(iy) == 1) x2 = -x2;
        *res = (iy >> 63) ? (1.0 / x2) : x2;
        return 1;
    }
    if (ix >> 63) {
        int yint = ode_pow_checkint_(iy);
        if (yint == 0) { *res = (x - x) / (x - x); return 1; } /* __math_invalid: NaN */
        if (yint == 1) sign_bias = ODE_POW_SIGN_BIAS;
        ix &= 0x7fffffffffffffffULL;
        topx &= 0x7ff;
    }
    if ((topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu) {
        if (ix == 0x3ff0000000000000ULL) { *res = 1.0; return 1; }
        if ((topy & 0x7ff) < 0x3beu) { /* |y| < 2^-65 */
            *res = ix > 0x3ff0000000000000ULL ? 1.0 + y : 1.0 - y;
            return 1;
        }
        *res = ((ix > 0x3ff0000000000000ULL) == (topy < 0x800u)) ? ODE_ASF64(0x7ff0000000000000ULL) : 0.0; /* __math_oflow(0) : __math_uflow(0) */
        return 1;
    }
    if (topx == 0) { /* subnormal x: normalise */
        ix = ODE_ASU64(x * 0x1p52);
        ix &= 0x7fffffffffffffffULL;
        ix -= 52ULL << 52;
    }
    *pix = ix;
    *psign_bias = sign_bias;
    return 0;
}

/* log(x) = *hi + *lo in double-double for ix = bits of a positive NORMAL x (log_inline). */
ODE_POW_FN void ode_pow_log_(ode_u64 ix, ode_pow_tabs T, double* phi, double* plo) {
    ode_u64 tmp = ix - ODE_POW_OFF;
    int i = (int)((tmp >> 45) & 127u);
    int k = (int)((ode_i64)tmp >> 52);
    ode_u64 iz = ix - (tmp & (0xfffULL << 52));
    double z = ODE_ASF64(iz);
    double kd = (double)k;
    double invc = ODE_ASF64(T.lt[3 * i]), logc = ODE_ASF64(T.lt[3 * i + 1]), logctail = ODE_ASF64(T.lt[3 * i + 2]);
    double r = fma(z, invc, -1.0);
    double t1 = fma(kd, K_LN2HI, logc);
    double t2 = t1 + r;
    double lo1 = fma(kd, K_LN2LO, logctail);
    double lo2 = t1 - t2 + r;
    double ar = K_A0 * r;
    double ar2 = r * ar;
    double ar3 = r * ar2;
    double hi = t2 + ar2;
    double lo3 = fma(ar, r, -ar2);
    double lo4 = t2 - hi + ar2;
    double q = fma(ar2, fma(r, K_A6, K_A5), fma(r, K_A4, K_A3));
    q = fma(ar2, q, fma(r, K_A2, K_A1));
    double lo = fma(ar3, q, lo1 + lo2 + lo3 + lo4);
    double y1 = hi + lo;
    *phi = y1;
    *plo = hi - y1 + lo;
}

/* exp(y * (hi + lo)) with the product in double-double: the second half of pow. Given the
 * log of one x it yields pow(x, y) for ANY y in pow's main range, so a caller that needs
 * pow(x, y1) and pow(x, y2) (the PI controller: err^alpha now, err^-beta for the next step)
 * pays for one log only -- same operations on the same values, hence the same bits. */
ODE_POW_FN double ode_pow_exp_(double y, double lhi, double llo, ode_u32 sign_bias, ode_pow_tabs T) {
    double ehi = y * lhi;
    double elo = fma(y, llo, fma(y, lhi, -ehi));
    return ode_exp_inline_(ehi, elo, sign_bias, 1, T);
}

/* true when pow(x, y) takes the main path for this x: positive, normal, finite */
ODE_POW_FN int ode_pow_x_is_plain_(double x) { return (ode_u32)(ODE_ASU64(x) >> 52) - 0x001u < 0x7ffu - 0x001u; }
/* true when pow(x, y) takes the main path for this y: 2^-65 <= |y| < 2^63 */
ODE_POW_FN int ode_pow_y_is_plain_(double y) {
    return ((ode_u32)(ODE_ASU64(y) >> 52) & 0x7ff) - 0x3beu < 0x43eu - 0x3beu;
}

/* pow(x, y), bit-identical to glibc 2.39 x86-64 (FMA variant). */
ODE_POW_FN double ode_pow(double x, double y, ode_pow_tabs T) {
    ode_u32 sign_bias = 0;
    ode_u64 ix = ODE_ASU64(x);
    ode_u32 topx = (ode_u32)(ix >> 52), topy = (ode_u32)(ODE_ASU64(y) >> 52);
    if (topx - 0x001u >= 0x7ffu - 0x001u || (topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu) {
        double res;
        if (ode_pow_special_(x, y, &ix, &sign_bias, &res)) return res;
    }
    double lhi, llo;
    ode_pow_log_(ix, T, &lhi, &llo);
    return ode_pow_exp_(y, lhi, llo, sign_bias, T);
}

/* exp(x), bit-identical to glibc 2.39 x86-64 (FMA variant). */
ODE_POW_FN double ode_exp(double x, ode_pow_tabs T) { return ode_exp_inline_(x, 0.0, 0, 0, T); }

#endif /* ODE_POW_H */
