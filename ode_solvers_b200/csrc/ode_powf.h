/*
 * ode_powf.h -- powf() that is BIT-IDENTICAL to the host's glibc 2.39 libm (x86-64, the FMA ifunc
 * variant), usable from C, C++ and CUDA: groundwork for the f32 instantiation of the adaptive
 * steppers (src/dop_shared.rs:74: `impl FloatNumber for f32`), whose PI controller calls f32::powf
 * twice per attempted step (src/controller.rs:53-54) and hinit once (src/dopri5.rs:236,
 * src/dop853.rs:244). See ode_pow.h for why "accurate" is not enough.
 *
 * Algorithm: the public table-driven powf of ARM Optimized Routines (as adopted by glibc >= 2.28),
 * evaluated in double: log2(x) from a 16-entry table + degree-5 polynomial, y*log2(x), 2^t from a
 * 32-entry table + degree-3 polynomial, rounded to float once. Tables: ode_powf_tables.h, read from
 * the running libm.so.6 (tools/extract_libm_powf_tables.py). Which multiply-adds are FUSED was read off
 * the machine code of glibc's FMA build (__powf_fma): every a*b+c of log2_inline and exp2_inline is one
 * vfmadd; r = xd - kd and the final y * s are a separate vsubsd / vmulsd.
 *
 * Upstream and licence: the algorithm is that of ARM Optimized Routines (math/pow.c, math/exp.c, math/powf.c;
 * MIT OR Apache-2.0 WITH LLVM-exception), which glibc ships under the LGPL-2.1-or-later in sysdeps/ieee754.
 * This file is a restatement written from the published algorithm and the observed machine code, not a copy of
 * either source; the numeric tables are data read at build time from the libm installed on the build host.
 *
 * tests/test_pow_clone.py checks bit equality with libm's powf on the host; the device build is the
 * same source compiled with -fmad=false (explicit fma() only).
 */
#ifndef ODE_POWF_H
#define ODE_POWF_H

#include "ode_powf_tables.h"

#if defined(__CUDACC__)
#define ODE_POWF_FN __host__ __device__ static __forceinline__
#else
#include <math.h>
#include <stdint.h>
#include <string.h>
#define ODE_POWF_FN static inline
#endif

#if defined(__CUDA_ARCH__)
#define ODE_F32_BITS(f) ((unsigned)__float_as_int(f))
#define ODE_F32_FROM(u) (__int_as_float((int)(u)))
#define ODE_F64_BITS(d) ((unsigned long long)__double_as_longlong(d))
#define ODE_F64_FROM(u) (__longlong_as_double((long long)(u)))
#else
ODE_POWF_FN unsigned ode_f32_bits_(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
ODE_POWF_FN float ode_f32_from_(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
ODE_POWF_FN unsigned long long ode_f64_bits_(double d) { unsigned long long u; memcpy(&u, &d, 8); return u; }
ODE_POWF_FN double ode_f64_from_(unsigned long long u) { double d; memcpy(&d, &u, 8); return d; }
#define ODE_F32_BITS(f) ode_f32_bits_(f)
#define ODE_F32_FROM(u) ode_f32_from_(u)
#define ODE_F64_BITS(d) ode_f64_bits_(d)
#define ODE_F64_FROM(u) ode_f64_from_(u)
#endif

/* lt = 16 x {invc, logc}, et = 32 exp2 table words, as raw 64-bit patterns (shared memory on the
 * device, static arrays on the host) */
typedef struct ode_powf_tabs {
    const unsigned long long* lt;
    const unsigned long long* et;
} ode_powf_tabs;

#if !defined(__CUDA_ARCH__)
static const unsigned long long ODE_POWF_LOG2_TAB_HOST[32] = ODE_POWF_LOG2_TAB_INIT;
static const unsigned long long ODE_EXP2F_TAB_HOST[32] = ODE_EXP2F_TAB_INIT;
ODE_POWF_FN ode_powf_tabs ode_powf_host_tabs(void) {
    ode_powf_tabs t;
    t.lt = ODE_POWF_LOG2_TAB_HOST;
    t.et = ODE_EXP2F_TAB_HOST;
    return t;
}
#endif

#define ODE_POWF_OFF 0x3f330000u
#define ODE_POWF_SIGN_BIAS (1u << 16) /* 1 << (EXP2F_TABLE_BITS + 11) */

/* 0 = not an integer, 1 = odd integer, 2 = even integer (iy: bits of a finite non-zero y) */
ODE_POWF_FN int ode_powf_checkint_(unsigned iy) {
    int e = (int)(iy >> 23) & 0xff;
    if (e < 0x7f) return 0;
    if (e > 0x7f + 23) return 2;
    if (iy & ((1u << (0x7f + 23 - e)) - 1u)) return 0;
    if (iy & (1u << (0x7f + 23 - e))) return 1;
    return 2;
}

ODE_POWF_FN int ode_powf_zeroinfnan_(unsigned i) { return 2u * i - 1u >= 2u * 0x7f800000u - 1u; }

/* log2(x) for ix = bits of a positive normal float (log2_inline) */
ODE_POWF_FN double ode_powf_log2_(unsigned ix, ode_powf_tabs T) {
    unsigned tmp = ix - ODE_POWF_OFF;
    int i = (int)((tmp >> (23 - 4)) % 16u);
    unsigned top = tmp & 0xff800000u;
    unsigned iz = ix - top;
    int k = (int)top >> 23; /* arithmetic shift */
    double invc = ODE_F64_FROM(T.lt[2 * i]), logc = ODE_F64_FROM(T.lt[2 * i + 1]);
    double z = (double)ODE_F32_FROM(iz);
    double r = fma(z, invc, -1.0);
    double y0 = logc + (double)k;
    double r2 = r * r;
    double y = fma(ODE_POWF_A0, r, ODE_POWF_A1);
    double p = fma(ODE_POWF_A2, r, ODE_POWF_A3);
    double r4 = r2 * r2;
    double q = fma(ODE_POWF_A4, r, y0);
    q = fma(p, r2, q);
    return fma(y, r4, q);
}

/* 2^xd * (-1)^(sign_bias != 0) rounded to float (exp2_inline); |xd| < 126 */
ODE_POWF_FN float ode_powf_exp2_(double xd, unsigned sign_bias, ode_powf_tabs T) {
    double kd = xd + ODE_EXP2F_SHIFT_SCALED;
    unsigned long long ki = ODE_F64_BITS(kd);
    kd -= ODE_EXP2F_SHIFT_SCALED;
    double r = xd - kd;
    unsigned long long t = T.et[ki % 32u];
    unsigned long long ski = ki + sign_bias;
    t += ski << (52 - 5);
    double s = ODE_F64_FROM(t);
    double z = fma(ODE_EXP2F_C0, r, ODE_EXP2F_C1);
    double r2 = r * r;
    double y = fma(ODE_EXP2F_C2, r, 1.0);
    y = fma(z, r2, y);
    y = y * s;
    return (float)y;
}

/* powf(x, y), bit-identical to glibc 2.39 x86-64 (FMA variant) in round-to-nearest */
ODE_POWF_FN float ode_powf(float x, float y, ode_powf_tabs T) {
    unsigned sign_bias = 0;
    unsigned ix = ODE_F32_BITS(x), iy = ODE_F32_BITS(y);
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u || ode_powf_zeroinfnan_(iy)) {
        /* either (x < 0x1p-126 or inf or nan) or (y is 0 or inf or nan) */
        if (ode_powf_zeroinfnan_(iy)) {
            if (2u * iy == 0u) return 1.0f; /* (a signalling NaN x also yields a quiet NaN / 1: payloads not compared) */
            if (ix == 0x3f800000u) return 1.0f;
            if (2u * ix > 2u * 0x7f800000u || 2u * iy > 2u * 0x7f800000u) return x + y;
            if (2u * ix == 2u * 0x3f800000u) return 1.0f;
            if ((2u * ix < 2u * 0x3f800000u) == !(iy & 0x80000000u)) return 0.0f; /* |x|<1 && y==inf or |x|>1 && y==-inf */
            return y * y;
        }
        if (ode_powf_zeroinfnan_(ix)) {
            float x2 = x * x;
            if ((ix & 0x80000000u) && ode_powf_checkint_(iy) == 1) x2 = -x2;
            return (iy & 0x80000000u) ? 1.0f / x2 : x2;
        }
        /* x and y are non-zero finite */
        if (ix & 0x80000000u) {
            int yint = ode_powf_checkint_(iy);
            if (yint == 0) return (x - x) / (x - x); /* __math_invalidf: NaN */
            if (yint == 1) sign_bias = ODE_POWF_SIGN_BIAS;
            ix &= 0x7fffffffu;
        }
        if (ix < 0x00800000u) { /* subnormal x: normalise */
            ix = ODE_F32_BITS(x * 0x1p23f);
            ix &= 0x7fffffffu;
            ix -= 23u << 23;
        }
    }
    double logx = ode_powf_log2_(ix, T);
    double ylogx = (double)y * logx; /* cannot overflow, y is single precision */
    if ((ODE_F64_BITS(ylogx) >> 47 & 0xffffu) >= (0x405f800000000000ULL >> 47)) { /* |y*log2(x)| >= 126 */
        if (ylogx > 0x1.fffffffd1d571p+6) { /* |x^y| > 0x1.ffffffp127: __math_oflowf */
            float big = sign_bias ? -0x1p97f : 0x1p97f;
            return big * 0x1p97f;
        }
        if (ylogx <= -150.0) { /* __math_uflowf */
            float tiny = sign_bias ? -0x1p-95f : 0x1p-95f;
            return tiny * 0x1p-95f;
        }
        if (ylogx < -149.0) { /* __math_may_uflowf */
            float tiny = sign_bias ? -0x1.4p-75f : 0x1.4p-75f;
            return tiny * 0x1.4p-75f;
        }
    }
    return ode_powf_exp2_(ylogx, sign_bias, T);
}

#endif /* ODE_POWF_H */
