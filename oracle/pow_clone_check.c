/*
 * pow_clone_check.c -- TEST INFRASTRUCTURE: compiles the product's ode_pow.h for the HOST so
 * tests can compare it bit-for-bit with the libm the reference crate's powf/exp resolve to.
 * Returns the number of mismatching elements (NaN == NaN counts as equal) and writes the
 * clone's results to `out` (may be NULL).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "../ode_solvers_b200/csrc/ode_pow.h"
#include "../ode_solvers_b200/csrc/ode_powf.h"

static int same_(double a, double b) {
    if (a != a && b != b) return 1;
    uint64_t ua, ub;
    memcpy(&ua, &a, 8);
    memcpy(&ub, &b, 8);
    return ua == ub;
}

int64_t ode_powcheck_pow(int64_t n, const double* x, const double* y, double* out, double* ref) {
    ode_pow_tabs T = ode_pow_host_tabs();
    int64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        double c = ode_pow(x[i], y[i], T), l = pow(x[i], y[i]);
        if (out) out[i] = c;
        if (ref) ref[i] = l;
        bad += !same_(c, l);
    }
    return bad;
}

int64_t ode_powcheck_exp(int64_t n, const double* x, double* out, double* ref) {
    ode_pow_tabs T = ode_pow_host_tabs();
    int64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        double c = ode_exp(x[i], T), l = exp(x[i]);
        if (out) out[i] = c;
        if (ref) ref[i] = l;
        bad += !same_(c, l);
    }
    return bad;
}

static int samef_(float a, float b) {
    if (a != a && b != b) return 1;
    uint32_t ua, ub;
    memcpy(&ua, &a, 4);
    memcpy(&ub, &b, 4);
    return ua == ub;
}

/* the f32 clone (ode_powf.h) against libm's powf */
int64_t ode_powcheck_powf(int64_t n, const float* x, const float* y, float* out, float* ref) {
    ode_powf_tabs T = ode_powf_host_tabs();
    int64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        float c = ode_powf(x[i], y[i], T), l = powf(x[i], y[i]);
        if (out) out[i] = c;
        if (ref) ref[i] = l;
        bad += !samef_(c, l);
    }
    return bad;
}
