// Stand-in for boost::math::spherical_harmonic_{r,i} (Boost is not available in the build container).
// TEST INFRASTRUCTURE ONLY: lets the reference's own external/Alm/Alm_cpp/activity.cpp compile where it lies, so that the
// product's tamcmc_host_alm can be pinned on the reference's Alm() (tests/test_oracle_vs_reference.py).
// Y_l^m(theta, phi) = sqrt((2l+1)/(4 pi) (l-m)!/(l+m)!) P_l^m(cos theta) e^{i m phi}  (Condon-Shortley phase; Boost's convention),
// with P_l^m from the standard upward recurrences in l -- generic in (l, m), not the closed forms the product tabulates.
#pragma once
#include <cmath>
namespace boost { namespace math {
namespace shim_detail {
inline long double assoc_legendre(int l, int m, long double x)      // m >= 0
{
    long double pmm = 1.0L;
    if (m > 0) {
        const long double somx2 = sqrtl((1.0L - x) * (1.0L + x));
        long double fact = 1.0L;
        for (int i = 1; i <= m; i++) { pmm *= -fact * somx2; fact += 2.0L; }
    }
    if (l == m) return pmm;
    long double pmmp1 = x * (2 * m + 1) * pmm;
    if (l == m + 1) return pmmp1;
    long double pll = 0.0L;
    for (int ll = m + 2; ll <= l; ll++) {
        pll = (x * (2 * ll - 1) * pmmp1 - (ll + m - 1) * pmm) / (ll - m);
        pmm = pmmp1; pmmp1 = pll;
    }
    return pll;
}
inline long double ylm_amplitude(int l, int m, long double theta)    // real amplitude of Y_l^m at phi = 0
{
    const int am = m < 0 ? -m : m;
    long double ratio = 1.0L;                                         // (l-am)!/(l+am)!
    for (int k = l - am + 1; k <= l + am; k++) ratio /= (long double)k;
    const long double pi = 3.141592653589793238462643383279502884L;
    long double a = sqrtl((2 * l + 1) / (4.0L * pi) * ratio) * assoc_legendre(l, am, cosl(theta));
    if (m < 0 && (am & 1)) a = -a;                                    // Y_l^{-m} = (-1)^m conj(Y_l^m)
    return a;
}
}  // namespace shim_detail
template <class T1, class T2> inline long double spherical_harmonic_r(unsigned l, int m, T1 theta, T2 phi)
{ return shim_detail::ylm_amplitude((int)l, m, (long double)theta) * cosl((long double)m * (long double)phi); }
template <class T1, class T2> inline long double spherical_harmonic_i(unsigned l, int m, T1 theta, T2 phi)
{ return shim_detail::ylm_amplitude((int)l, m, (long double)theta) * sinl((long double)m * (long double)phi); }
}}  // namespace boost::math
