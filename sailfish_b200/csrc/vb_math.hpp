// vb_math.hpp -- digamma and exp(digamma) as VBEM uses them (CollapsedEMOptimizer.cpp:398-416).  CUDA-free: compiled by nvcc into the
// EM kernels (common.cuh) and by g++ into the CPU test of the same functions (tests/vb_math_test.cpp).
#pragma once
#include <cmath>
#if defined(__CUDACC__)
#define SFB_VB_HD __host__ __device__
#else
#define SFB_VB_HD
#endif

// digamma for x > 0: recurrence up to x >= 12, then the asymptotic series (same expansion as the oracle's
// stand-in for boost::math::digamma; checked against scipy in tests/).  The recurrence psi(x) = psi(x + n) - sum 1/(x + k) is the
// expensive part of a VBEM iteration (up to twelve fp64 divisions per transcript): the sum of reciprocals is P'/P of the polynomial
// P = prod (x + k), built with two multiplies and an FMA per factor and ONE division -- after a first term taken by itself
// when x < 1, so that P stays far from the denormal range for the tiny alphas VBEM produces.
SFB_VB_HD inline double sfb_digamma(double x) {
    double acc = 0.0;
    if (x < 1.0) { acc = -1.0 / x; x += 1.0; }
    if (x < 12.0) {
        double P = 1.0, Q = 0.0;                     // P = prod (x + k), Q = dP/dx
        while (x < 12.0) { Q = fma(Q, x, P); P *= x; x += 1.0; }
        acc -= Q / P;
    }
    const double inv = 1.0 / x, inv2 = inv * inv;
    const double series = inv2 * (1.0 / 12.0 - inv2 * (1.0 / 120.0 - inv2 * (1.0 / 252.0 - inv2 * (1.0 / 240.0 -
                          inv2 * (1.0 / 132.0 - inv2 * (691.0 / 32760.0 - inv2 * (1.0 / 12.0)))))));
    return acc + log(x) - 0.5 * inv - series;
}

// exp(digamma(x)) for x > 0 without the logarithm: VBEM's expTheta is exp(digamma(alpha) - digamma(sum alpha))
// (CollapsedEMOptimizer.cpp:398-416) = sfb_exp_digamma(alpha) * exp(-digamma(sum alpha)), the second factor one number per
// iteration.  For x >= 16 the asymptotic series exp(psi(x)) = x (1 - y/2 + y^2/24 + y^3/48 + 23 y^4/5760 - ...), y = 1/x (the
// exponential of psi's own series, coefficients exact rationals), is good to 2e-16 with twelve terms and costs one division and
// twelve FMAs where exp(log(x) - ...) cost a logarithm, an exponential and a division; below 16 the recurrence of sfb_digamma
// shifts x up and its sum of reciprocals goes through ONE exponential.  Relative error against mpmath over [1e-3, 1e9]: 1.1e-13
// (from exp of the -1/x term of tiny alphas), the exp(digamma - logNorm) form it replaces: 1.8e-13 (tests/test_vb_math.py runs
// the host build of both against mpmath; tests/test_gpu_em.py the device build).
SFB_VB_HD inline double sfb_exp_digamma(double x) {
    double mult = 1.0;
    if (x < 16.0) {
        double acc = 0.0;
        if (x < 1.0) { acc = -1.0 / x; x += 1.0; }
        double P = 1.0, Q = 0.0;
        while (x < 16.0) { Q = fma(Q, x, P); P *= x; x += 1.0; }
        acc -= Q / P;
        mult = exp(acc);
    }
    const double y = 1.0 / x;
    double s = 318246113.0 / 81749606400.0;
    s = fma(s, y, -2727899759.0 / 367873228800.0);
    s = fma(s, y, -870041.0 / 398131200.0);
    s = fma(s, y, 795697.0 / 199065600.0);
    s = fma(s, y, 2501.0 / 1161216.0);
    s = fma(s, y, -10099.0 / 2903040.0);
    s = fma(s, y, -17.0 / 3840.0);
    s = fma(s, y, 23.0 / 5760.0);
    s = fma(s, y, 1.0 / 48.0);
    s = fma(s, y, 1.0 / 24.0);
    s = fma(s, y, -0.5);
    s = fma(s, y, 1.0);
    return x * s * mult;
}

// VBEM's expTheta = exp(digamma(alpha) - logNorm), logNorm = digamma(sum alpha) -- one number per iteration, and so is
// scale = exp(-logNorm).  The product form needs scale to be an ordinary number (sum alpha >= 1 gives logNorm >= -0.58; the other
// form stays for a degenerate sum).
SFB_VB_HD inline double sfb_exp_theta(double alpha, double logNorm, double scale) {
    return (logNorm > -600.0 && logNorm < 600.0) ? sfb_exp_digamma(alpha) * scale : exp(sfb_digamma(alpha) - logNorm);
}
