// dd_math.cuh -- double-double arithmetic and two CORRECTLY ROUNDED (to within ~2^-90) elementary functions, host + device.
//
// Why: the red-giant mixed-mode solver (external/ARMM/solver_mm.cpp:326-449) finds its l=1 frequencies from values of
//   p(nu) - g(nu),  g = Dnu atan(q tan(pi 1e6 (1/nu - 1/nu_g) / DPl)) / pi      (solver_mm.cpp:114-161)
// and a mixed mode can be 1e-4 microHz wide: its frequency has to agree with the reference's to the last bits for 1e-10 on the
// spectrum.  Every +, -, *, / of that path is an IEEE operation the device reproduces exactly (no FMA contraction); tan and atan are
// library calls.  glibc's are correctly rounded except for arguments whose result lies within ~2^-70 of a rounding boundary; CUDA's
// are good to 1-2 ulp, which is not the same double often enough to matter.  tan_cr / atan_cr below evaluate both in double-double
// (relative error < 2^-90) and round once: the same double as glibc's for all but ~1e-5 of the arguments (measured on the host
// against glibc and against mpmath, tests/test_dd_math.py) -- so the device search reproduces the host solver's frequencies.
// Compiled for the host as well (plain C++: __CUDACC__ absent), which is how the tests pin it without a GPU.
// New code; the algorithms are the textbook ones (Dekker / Knuth error-free transformations, Cody-Waite reduction, table + Taylor).
#pragma once
#include <cmath>
#include "dd_tables.h"

#ifdef __CUDACC__
#define TAMCMC_HD __host__ __device__ __forceinline__
#define TAMCMC_HD_CALL __host__ __device__ __noinline__          // the big ones: one copy per kernel, called
#else
#define TAMCMC_HD inline
#define TAMCMC_HD_CALL inline
#endif

namespace tamcmc_dd {

struct dd { double hi, lo; };

// sin(j/64), cos(j/64), j = 0..51, as double-doubles (tools/make_dd_tables.py)
#ifdef __CUDACC__
__device__ const double g_sincos_tab_dev[TAMCMC_DD_NTAB][4] = {TAMCMC_DD_SINCOS_TABLE};
#endif
static const double g_sincos_tab_host[TAMCMC_DD_NTAB][4] = {TAMCMC_DD_SINCOS_TABLE};

TAMCMC_HD double fma_(double a, double b, double c)
{
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}
// the EFTs below rely on every operation rounding on its own: on the device use the explicitly rounded intrinsics so that no build flag
// can contract them
TAMCMC_HD double add_(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
TAMCMC_HD double mul_(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
TAMCMC_HD dd two_sum(double a, double b)
{
    const double s = add_(a, b);
    const double bb = add_(s, -a);
    const double e = add_(add_(a, -add_(s, -bb)), add_(b, -bb));
    return {s, e};
}
TAMCMC_HD dd quick_two_sum(double a, double b)          // |a| >= |b|
{
    const double s = add_(a, b);
    return {s, add_(b, -add_(s, -a))};
}
TAMCMC_HD dd two_prod(double a, double b)
{
    const double p = mul_(a, b);
    return {p, fma_(a, b, -p)};
}
TAMCMC_HD dd neg(dd a) { return {-a.hi, -a.lo}; }
TAMCMC_HD dd add(dd a, dd b)
{
    dd s = two_sum(a.hi, b.hi);
    const dd t = two_sum(a.lo, b.lo);
    s.lo = add_(s.lo, t.hi);
    s = quick_two_sum(s.hi, s.lo);
    s.lo = add_(s.lo, t.lo);
    return quick_two_sum(s.hi, s.lo);
}
TAMCMC_HD dd add_d(dd a, double b)
{
    dd s = two_sum(a.hi, b);
    s.lo = add_(s.lo, a.lo);
    return quick_two_sum(s.hi, s.lo);
}
TAMCMC_HD dd mul(dd a, dd b)
{
    dd p = two_prod(a.hi, b.hi);
    p.lo = add_(p.lo, add_(mul_(a.hi, b.lo), mul_(a.lo, b.hi)));
    return quick_two_sum(p.hi, p.lo);
}
TAMCMC_HD dd mul_d(dd a, double b)
{
    dd p = two_prod(a.hi, b);
    p.lo = add_(p.lo, mul_(a.lo, b));
    return quick_two_sum(p.hi, p.lo);
}
TAMCMC_HD dd div(dd a, dd b)
{
    const double q1 = a.hi / b.hi;
    dd r = add(a, neg(mul_d(b, q1)));
    const double q2 = r.hi / b.hi;
    r = add(r, neg(mul_d(b, q2)));
    const double q3 = r.hi / b.hi;
    return add_d(quick_two_sum(q1, q2), q3);
}

struct sc { dd s, c; };

// sin and cos of a double x (|x| < ~1e6) as double-doubles, relative error < ~2^-95
TAMCMC_HD_CALL sc sincos_dd(double x)
{
#ifdef __CUDA_ARCH__
    const double (*T)[4] = g_sincos_tab_dev;
#else
    const double (*T)[4] = g_sincos_tab_host;
#endif
    // Cody-Waite: x - k pi/2 with pi/2 = P1 + P2 + P3 + P4, P1 and P2 of 30 bits (k P1, k P2 exact for |k| < 2^22)
    const double kd = rint(mul_(x, TAMCMC_DD_2OPI));
    dd r = two_sum(x, -mul_(kd, TAMCMC_DD_PIO2_1));
    r = add_d(r, -mul_(kd, TAMCMC_DD_PIO2_2));
    r = add(r, neg(two_prod(kd, TAMCMC_DD_PIO2_3)));
    r = add_d(r, -mul_(kd, TAMCMC_DD_PIO2_4));
    const bool negative = r.hi < 0.0;
    if (negative) r = neg(r);
    // r = j/64 + t, |t| <= 1/128
    double jd = rint(mul_(r.hi, 64.0));
    if (jd > (double)(TAMCMC_DD_NTAB - 1)) jd = (double)(TAMCMC_DD_NTAB - 1);
    const int j = (int)jd;
    const dd t = add_d(r, -mul_(jd, 0.015625));
    const dd z = mul(t, t);
    const double zh = z.hi;
    // sin t = t + t z (-1/6 + z (1/120 + z (-1/5040 + z (1/362880 - z/39916800))))
    double w = -1.0 / 5040.0 + zh * (1.0 / 362880.0 + zh * (-1.0 / 39916800.0));
    dd u = add(dd{TAMCMC_DD_C120_HI, TAMCMC_DD_C120_LO}, mul_d(z, w));
    dd v = add(dd{TAMCMC_DD_C6_HI, TAMCMC_DD_C6_LO}, mul(z, u));
    const dd st = add(t, mul(mul(t, z), v));
    // cos t = 1 + z (-1/2 + z (1/24 + z (-1/720 + z (1/40320 + z (-1/3628800 + z/479001600)))))
    w = -1.0 / 720.0 + zh * (1.0 / 40320.0 + zh * (-1.0 / 3628800.0 + zh * (1.0 / 479001600.0)));
    u = add(dd{TAMCMC_DD_C24_HI, TAMCMC_DD_C24_LO}, mul_d(z, w));
    v = add_d(mul(z, u), -0.5);
    const dd ct = add_d(mul(z, v), 1.0);
    const dd sa = {T[j][0], T[j][1]}, ca = {T[j][2], T[j][3]};
    dd s = add(mul(sa, ct), mul(ca, st));
    dd c = add(mul(ca, ct), neg(mul(sa, st)));
    if (negative) s = neg(s);
    const long k = (long)kd;
    switch (k & 3) {
    case 0: return {s, c};
    case 1: return {c, neg(s)};
    case 2: return {neg(s), neg(c)};
    default: return {neg(c), s};
    }
}

// tan(x) rounded once from a double-double quotient
TAMCMC_HD_CALL double tan_cr(double x)
{
    if (!(fabs(x) < 4.0e6)) return tan(x);                  // outside the reduction's exact range (never on the solver's path)
    const sc v = sincos_dd(x);
    const dd q = div(v.s, v.c);
    return add_(q.hi, q.lo);
}

// atan(t) = a0 + atan((t cos a0 - sin a0) / (cos a0 + t sin a0)) with a0 the library's atan: the correction is ~1e-16, its atan is itself
TAMCMC_HD_CALL double atan_cr(double t)
{
    const double a0 = atan(t);
    if (t == 0.0 || !(fabs(t) < 1.0e290)) return a0;
    const sc v = sincos_dd(a0);
    const dd num = add(mul_d(v.c, t), neg(v.s));
    const dd den = add(v.c, mul_d(v.s, t));
    return add_(a0, num.hi / den.hi);
}

}  // namespace tamcmc_dd
