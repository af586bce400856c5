// rgb_solver.cuh -- the pair loop of the red-giant mixed-mode solver as independent SEGMENT tasks (host + device).
//
// Replaces (on the device): solver_mm of external/ARMM/solver_mm.cpp:326-449 as host_rgb.cpp restates it -- for one (p mode, g mode)
// pair, the sign changes of f(nu) = p(nu) - g(nu) on the coarse grid of the p mode's band and, for each, the solution refined on the
// local grid by the reference's lin_interpol (tamcmc/sources/interpol.cpp:13-43) and kept if g / p is within 0.1 % of 1.
//
// Decomposition.  g = Dnu atan(q tan(X)) / pi, X = pi 1e6 (1/nu - 1/nu_g) / DPl, jumps from +Dnu/2 to -Dnu/2 ... i.e. f jumps DOWN at
// every pole of the tangent and rises with slope > 1 in between: between two poles f changes sign at most once (- to +), at a pole it
// may change from + to -.  The poles are known analytically (X / pi = m + 1/2), so segment j of a pair = the stretch in front of
// pole j (one bisection on the grid index) + the four grid points around pole j (compared pair by pair like sign_change,
// solver_mm.cpp:71-106).  Segments are independent: one thread each, no communication; their solutions are appended to the chain's
// candidate list, which the host filters, sorts and makes unique exactly like the reference (solver_mm.cpp:575-590).
//
// Two phases.  The g modes of one p mode see the SAME function up to rounding (tan has period pi and 1/nu_g = (ng + alpha) DPl 1e-6): the same
// sign changes at the same grid indices, each with its own last bits -- and the reference keeps the smallest solution of every cluster,
// so every (p mode, g mode) version has to be evaluated, but the SEARCH is done once: phase 1 runs the segments of a band with the
// band's first g mode and leaves one record per sign change (coarse index, local grid, which two local points the interpolation
// uses); phase 2 evaluates every record for every g mode of the band -- two exact values of f, one line, one ratio test: uniform work,
// no divergence.  Phase 1 flags the chain when a value it decides on is smaller than 1e-11 (another g mode's rounding could decide
// otherwise); phase 2 checks the signs of its two values against the record.
//
// Arithmetic.  Every value that enters the interpolation is computed with the operations of the reference in its order (no FMA
// contraction: this header is compiled with -fmad=false / -ffp-contract=off) and the correctly rounded tan / atan of dd_math.cuh;
// searches only need the SIGN of f and use the fast library functions, re-evaluated exactly when |f| < 1e-9.  Anything the
// decomposition does not expect (an exact zero, a stretch that does not rise, two poles in one local grid, a ratio test within 1e-6
// of its bounds) raises the chain's flag and the host solves that chain.
#pragma once
#include "dd_math.cuh"
#include "host_rgb.hpp"

namespace tamcmc_rgb {

struct TrigCR {
    static TAMCMC_HD double tan_(double x) { return tamcmc_dd::tan_cr(x); }
    static TAMCMC_HD double atan_(double x) { return tamcmc_dd::atan_cr(x); }
};
struct TrigLib {
    static TAMCMC_HD double tan_(double x) { return tan(x); }
    static TAMCMC_HD double atan_(double x) { return atan(x); }
};

#define TAMCMC_RGB_PI 3.141592653589793

struct Record { double lo, hi; int idx, n, kind, i; };     // kind 0: f(i) < 0 < f(i+1) on the local grid; 1: line through points 0, 1; 2: through n-2, n-1

enum { RGB_FLAG_NEAR = 256, RGB_FLAG_VERIFY = 512, RGB_FLAG_ZERO = 1, RGB_FLAG_SHAPE = 2, RGB_FLAG_POLES = 4, RGB_FLAG_RATIO = 8, RGB_FLAG_OVERFLOW = 16, RGB_FLAG_NONFINITE = 32, RGB_FLAG_EXT = 64 };


// p(nu) - g(nu) as the VectorXd versions of pnu_fct / gnu_fct compute one element (solver_mm.cpp:114-161; host_rgb.cpp PminusG)
template <class TR>
TAMCMC_HD_CALL double pmg(const Band& B, double inv_g, double nu)
{
    const double pnu = nu - B.nu_p;
    const double X = ((TAMCMC_RGB_PI * (1.0 / nu - inv_g)) * 1e6) / B.DPl;
    const double t = B.q * TR::tan_(X);
    const double gnu = (B.Dnu * TR::atan_(t)) / TAMCMC_RGB_PI;
    return pnu - gnu;
}
// a value whose sign is the sign of the exact value -- for EVERY g mode of the band, or the chain is flagged
template <class FAST, class EXACT>
TAMCMC_HD double pmg_sign(const Band& B, double inv_g, double nu, int& flag)
{
    double v = pmg<FAST>(B, inv_g, nu);
    if (!(fabs(v) > 1e-9)) {
        v = pmg<EXACT>(B, inv_g, nu);
        if (!(fabs(v) > 1e-11)) flag |= RGB_FLAG_NEAR;
    }
    return v;
}
TAMCMC_HD double u_of(const Band& B, double inv_g, double nu) { return (1.0 / nu - inv_g) * 1e6 / B.DPl; }
TAMCMC_HD double nu_of_u(const Band& B, double inv_g, double u) { return 1.0 / (u * B.DPl / 1e6 + inv_g); }
TAMCMC_HD double band_nu(const Band& B, int i)                     // Eigen::VectorXd::LinSpaced(n, lo, hi)[i_lo + i]
{
    const int k = B.i_lo + i;
    return (k == B.n - 1) ? B.hi : B.lo + (double)k * B.gstep;
}

// ---- the local grid of a sign change, as the reference derives it in long double (solver_mm.cpp:402-410; host_rgb.cpp local_grid):
//   range_min = nu - 2 resol, range_max = nu + 2 resol                       (x87 extended: 64-bit mantissa)
//   nu_local  = linspaced((long)((range_max - range_min) / (resol * factor)), (double)range_min, (double)range_max)
// reproduced with error-free transformations.  a + b = s + e exactly (two_sum); rounding to 64 bits rounds e to a multiple of
// ulp(s) / 2048 (ties to even: the parity is that of e's multiple because s is an even multiple), and the conversion to double gives s
// unless the rounded e is exactly half an ulp of s (then the even neighbour).  The difference of the two extended values is exact; the
// quotient by D = resol * factor (an extended value the host passes as two doubles) is compared with the nearest integer k through the
// exact residual N - k D: the truncated extended quotient is k when N / D >= k - (half an extended ulp below k), else k - 1.
// false (and the chain goes to the host) at a binade boundary, where the spacings above do not hold.
struct ExtSum { double d, r; bool ok; };                 // the extended value is d + r, d its conversion to double
TAMCMC_HD ExtSum ext_add(double a, double b)
{
    using namespace tamcmc_dd;
    const dd se = two_sum(a, b);
    const double s = se.hi, e = se.lo;
    int ex;
    const double fr = frexp(fabs(s), &ex);                    // |s| = fr 2^ex, fr in [0.5, 1)
    ExtSum R;
    R.ok = (fr != 0.5) && (s == s) && fabs(s) > 1e-290 && fabs(s) < 1e290;
    const double u = ldexp(1.0, ex - 53);                     // ulp of s
    const double M = ldexp(1.5, ex - 53 - 11 + 52);           // (e + M) - M rounds e to a multiple of u / 2048, ties to even
    const double er = add_(add_(e, M), -M);
    R.d = s; R.r = er;
    if (fabs(er) == 0.5 * u) {                                // a tie of the second rounding: to the even double
        const double t = s / u;                               // an integer below 2^53
        const bool odd = fmod(t, 2.0) != 0.0;
        if (odd) { R.d = (er > 0.0) ? s + u : s - u; R.r = (er > 0.0) ? -0.5 * u : 0.5 * u; }
    }
    return R;
}
TAMCMC_HD bool local_grid_ext(double nu, double resol2 /* 2 resol */, double Dh, double Dl /* resol * factor, extended */, double& lo, double& hi, int& n)
{
    using namespace tamcmc_dd;
    const ExtSum A = ext_add(nu, -resol2), B = ext_add(nu, resol2);
    lo = A.d; hi = B.d;
    if (!A.ok || !B.ok || !(Dh > 0.0)) return false;
    // N = range_max - range_min exactly: (B.d - A.d) + (B.r - A.r), both differences exact
    const dd N = two_sum(B.d - A.d, B.r - A.r);
    const double q0 = N.hi / Dh;
    if (!(q0 == q0) || q0 >= 1.0e9) return false;
    if (q0 < 1.5) { n = 0; return true; }                     // fewer than two points: the reference skips the sign change
    const double k = rint(q0);
    if (fabs(q0 - k) > 1e-6) { n = (int)q0; return true; }    // not next to an integer: the truncation cannot depend on the last bits
    const dd P = two_prod(k, Dh);                             // k D = P.hi + P.lo + k Dl, every piece exact
    const double rho = add_(add_(N.hi, -P.hi), add_(add_(N.lo, -P.lo), -mul_(k, Dl)));
    int ek;
    frexp(k - 0.5, &ek);                                      // k - 0.5 in [2^(ek-1), 2^ek): extended ulp there is 2^(ek-64)
    const double h = ldexp(1.0, ek - 65);                     // half of it
    n = (rho >= -h * Dh) ? (int)k : (int)k - 1;
    return true;
}


// phase 1: which two points of linspaced(Nx, lo, hi) does lin_interpol(f(nu_local), nu_local, 0) (tamcmc/sources/interpol.cpp:13-43; host_rgb.cpp
// interp_zero_lazy) draw its line through?  false: nothing to evaluate (or the chain is flagged)
template <class FAST, class EXACT>
TAMCMC_HD bool local_search(const Band& B, double inv_g, double lo, double hi, int Nx, int& kind, int& ibr, int& flag)
{
    if (Nx < 2) return false;
    const double lstep = (hi - lo) / (double)(Nx - 1);
    auto y = [&](int i) { return (i == Nx - 1) ? hi : lo + (double)i * lstep; };
    auto S = [&](int i) { return pmg_sign<FAST, EXACT>(B, inv_g, y(i), flag); };
    const double X0 = S(0), XN = S(Nx - 1);
    if (!(X0 == X0) || !(XN == XN)) { flag |= RGB_FLAG_NONFINITE; return false; }
    if (X0 == 0.0 || XN == 0.0) { flag |= RGB_FLAG_ZERO; return false; }
    if (0.0 > XN) { kind = 2; ibr = Nx - 2; return true; }                    // the last assignment of lin_interpol wins
    if (0.0 < X0) { kind = 1; ibr = 0; return true; }
    // X0 < 0 < XN: the first i with f(i) <= 0 <= f(i+1); f rises except for a jump down at a pole of the tangent
    const double uA = u_of(B, inv_g, y(0)), uB = u_of(B, inv_g, y(Nx - 1));          // uA > uB
    const double mA = floor(uA - 0.5), mB = ceil(uB - 0.5);
    if (!(mA - mB < 1.0)) { flag |= RGB_FLAG_POLES; return false; }                  // two poles (or not finite): the host's walk
    int pts[7], np = 0;
    pts[np++] = 0;
    if (mA >= mB) {
        const double nup = nu_of_u(B, inv_g, mA + 0.5);
        const int il = (int)floor((nup - lo) / lstep);
        for (int k = il - 1; k <= il + 2; k++) if (k > pts[np - 1] && k <= Nx - 1) pts[np++] = k;
    }
    if (Nx - 1 > pts[np - 1]) pts[np++] = Nx - 1;
    int i = Nx - 2;
    bool found = false;
    double fa = X0;
    for (int s = 0; s + 1 < np && !found; s++) {
        const int pa = pts[s], pb = pts[s + 1];
        const double fb = (pb == Nx - 1) ? XN : S(pb);
        if (!(fb == fb) || fb == 0.0) { flag |= (fb == 0.0) ? RGB_FLAG_ZERO : RGB_FLAG_NONFINITE; return false; }
        if (pb == pa + 1) {
            if (fa < 0.0 && fb > 0.0) { i = pa; found = true; }
        } else if (fa < 0.0 && fb > 0.0) {
            int l = pa, h = pb;
            while (h - l > 1) {
                const int mid = l + (h - l) / 2;
                const double v = S(mid);
                if (!(v == v) || v == 0.0) { flag |= (v == 0.0) ? RGB_FLAG_ZERO : RGB_FLAG_NONFINITE; return false; }
                if (v < 0.0) l = mid; else h = mid;
            }
            i = l; found = true;
        } else if (fa > 0.0 && fb < 0.0) { flag |= RGB_FLAG_SHAPE; return false; }     // a stretch without a pole does not fall
        fa = fb;
    }
    if (!found) { flag |= RGB_FLAG_SHAPE; return false; }
    if (i > Nx - 2) i = Nx - 2;
    kind = 0; ibr = i;
    return true;
}

// g(nu) / p(nu) of a proposed solution, as the reference forms it in long double (gnu_fct for one frequency, solver_mm.cpp:172-180, then
// :421-427), evaluated in double-double: X = pi 1e6 (1/nu - 1/nu_g) / DPl with its low part carried through the tangent to first order,
// atan as the library value plus its double-double correction.  Good to ~1e-30; the reference's own extended-precision value differs
// from it by its rounding errors, a few 1e-15 / |p| on the ratio at most.
TAMCMC_HD_CALL double ratio_dd(const Band& B, double nu_g, double nu_m)
{
    using namespace tamcmc_dd;
    const dd pi = {3.141592653589793, 1.2246467991473532e-16};
    const dd one = {1.0, 0.0};
    dd X = add(div(one, dd{nu_m, 0.0}), neg(div(one, dd{nu_g, 0.0})));
    X = div(mul_d(mul(pi, X), 1e6), dd{B.DPl, 0.0});
    const sc v = sincos_dd(X.hi);
    dd t = div(v.s, v.c);
    t = add(t, mul_d(add(one, mul(t, t)), X.lo));                       // tan(X.hi + X.lo)
    const dd u = mul_d(t, B.q);
    const double a0 = atan(u.hi);
    const sc w = sincos_dd(a0);
    const dd num = add(mul(w.c, u), neg(w.s)), den = add(w.c, mul(w.s, u));
    const dd at = add_d(div(num, den), a0);                             // atan(u) = a0 + atan((u cos a0 - sin a0) / (cos a0 + u sin a0))
    const dd g = div(mul_d(at, B.Dnu), pi);
    const dd r = div(g, dd{nu_m - B.nu_p, 0.0});
    return r.hi + r.lo;
}

// phase 2: the line through the record's two local points with THIS g mode's exact values of f, its zero, and the 0.1 % test
// (solver_mm.cpp:421-431).  The signs of the two values must be what the record's kind says they are.
template <class EXACT>
TAMCMC_HD bool record_eval(const Band& B, double inv_g, double nu_g, const Record& R, double& sol, int& flag)
{
    const int Nx = R.n;
    const double lo = R.lo, hi = R.hi;
    const double lstep = (hi - lo) / (double)(Nx - 1);
    auto y = [&](int i) { return (i == Nx - 1) ? hi : lo + (double)i * lstep; };
    const int i = R.i;
    const double Xa = pmg<EXACT>(B, inv_g, y(i)), Xb = pmg<EXACT>(B, inv_g, y(i + 1));
    double a, b;
    if (R.kind == 0) {
        if (!(Xa < 0.0 && Xb > 0.0)) { flag |= RGB_FLAG_VERIFY; return false; }
        a = (y(i + 1) - y(i)) / (Xb - Xa); b = y(i) - a * Xa;
    } else if (R.kind == 1) {
        if (!(Xa > 0.0)) { flag |= RGB_FLAG_VERIFY; return false; }
        a = (y(1) - y(0)) / (Xb - Xa); b = y(0) - a * Xa;
    } else {
        if (!(Xb < 0.0)) { flag |= RGB_FLAG_VERIFY; return false; }
        a = (y(Nx - 1) - y(Nx - 2)) / (Xb - Xa); b = y(Nx - 2) - a * Xa;
    }
    const double x_int = 0.0;
    const double nu_m = a * x_int + b;
    // g / p within 0.1 % of 1.  The reference evaluates g in long double: plain double decides unless the ratio is within 1e-6 of a bound,
    // then the double-double value does, unless THAT is within the reach of the reference's own rounding errors (the chain goes to the host)
    const double Xs = TAMCMC_RGB_PI * (1.0 / nu_m - inv_g) * 1e6 / B.DPl;
    double ratio = (B.Dnu * atan(B.q * tan(Xs)) / TAMCMC_RGB_PI) / (nu_m - B.nu_p);
    if (fabs(ratio - 0.999) < 1e-6 || fabs(ratio - 1.001) < 1e-6) {
        ratio = ratio_dd(B, nu_g, nu_m);
        const double band = 4e-14 + 1.6e-13 / fabs(nu_m - B.nu_p);       // 4 x what 3e6 random cases need (tests/cpp/ratio_dd_check.cpp)
        if (!(fabs(ratio - 0.999) > band && fabs(ratio - 1.001) > band)) { flag |= RGB_FLAG_RATIO; return false; }
    }
    if (ratio >= 0.999 && ratio <= 1.001) { sol = nu_m; return true; }
    return false;
}

// number of segments of a pair (poles of the tangent inside the band + 1); 0 and a flag when the decomposition does not apply
TAMCMC_HD int pair_segments(const Band& B, double inv_g, double& m_hi, double& nu0, double& bstep, int& flag)
{
    const int nb = B.nband;
    nu0 = band_nu(B, 0);
    const double nuN = band_nu(B, nb - 1);
    bstep = (nuN - nu0) / (double)(nb - 1);
    if (!(bstep > 0.0) || !(nu0 > 0.0)) { flag |= RGB_FLAG_SHAPE; return 0; }
    const double u_hi = u_of(B, inv_g, nu0), u_lo = u_of(B, inv_g, nuN);
    m_hi = floor(u_hi - 0.5);
    const double m_lo = ceil(u_lo - 0.5);
    if (!(m_hi - m_lo <= (double)nb / 6.0) || !(m_hi - m_lo > -4.0)) { flag |= RGB_FLAG_POLES; return 0; }     // poles closer than ~6 grid steps (or NaN): the host's scan
    const int npoles = (m_hi >= m_lo) ? (int)(m_hi - m_lo) + 1 : 0;
    return npoles + 1;
}

// phase 1, segment j of a band (with the band's first g mode): emit(record) for every sign change that has something to evaluate
template <class FAST, class EXACT, class Emit>
TAMCMC_HD void pair_segment(const Band& B, double inv_g, int j, int nseg, double m_hi, double nu0, double bstep, Emit&& emit, int& flag)
{
    const int nb = B.nband, npoles = nseg - 1;
    auto ip_of = [&](int k) { return (int)floor((nu_of_u(B, inv_g, (m_hi - (double)k) + 0.5) - nu0) / bstep); };
    auto S = [&](int i) { return pmg_sign<FAST, EXACT>(B, inv_g, band_nu(B, i), flag); };
    int pts[7], np = 0;
    int start = 0;
    if (j > 0) { start = ip_of(j - 1) + 2; if (start < 0) start = 0; if (start > nb - 1) start = nb - 1; }
    pts[np++] = start;
    if (j < npoles) {
        const int ip = ip_of(j);
        for (int k = ip - 1; k <= ip + 2; k++) if (k > pts[np - 1] && k <= nb - 1) pts[np++] = k;
    } else if (nb - 1 > pts[np - 1]) pts[np++] = nb - 1;
    double fa = S(pts[0]);
    if (!(fa == fa) || fa == 0.0 || fabs(fa) > 1.0e300) { flag |= (fa == 0.0) ? RGB_FLAG_ZERO : RGB_FLAG_NONFINITE; return; }
    for (int s = 0; s + 1 < np; s++) {
        const int pa = pts[s], pb = pts[s + 1];
        const double fb = S(pb);
        if (!(fb == fb) || fb == 0.0 || fabs(fb) > 1.0e300) { flag |= (fb == 0.0) ? RGB_FLAG_ZERO : RGB_FLAG_NONFINITE; return; }
        int idx = -1;
        if (pb == pa + 1) {
            if ((fb > 0.0 && fa < 0.0) || (fb < 0.0 && fa > 0.0)) idx = pa;            // the two tests of sign_change; no value is zero here
        } else {
            if (fa > 0.0 && fb < 0.0) { flag |= RGB_FLAG_SHAPE; return; }
            if (fa < 0.0 && fb > 0.0) {
                int l = pa, h = pb;
                while (h - l > 1) {
                    const int mid = l + (h - l) / 2;
                    const double v = S(mid);
                    if (!(v == v) || v == 0.0 || fabs(v) > 1.0e300) { flag |= (v == 0.0) ? RGB_FLAG_ZERO : RGB_FLAG_NONFINITE; return; }
                    if (v < 0.0) l = mid; else h = mid;
                }
                idx = l;
            }
        }
        if (idx >= 0) {
            Record R;
            R.idx = idx;
            if (!local_grid_ext(band_nu(B, idx), B.resol2, B.Dh, B.Dl, R.lo, R.hi, R.n)) { flag |= RGB_FLAG_EXT; return; }
            if (local_search<FAST, EXACT>(B, inv_g, R.lo, R.hi, R.n, R.kind, R.i, flag)) emit(R);
        }
        fa = fb;
    }
}

}  // namespace tamcmc_rgb
