// host_rgb.cpp -- host expander of the red-giant models (BASELINE configs C1 / C4): the host half of
//   model_RGB_asympt_aj_AppWidth_HarveyLike_v4   tamcmc/sources/models.cpp:4684-5079   (models_ctrl.list id 25)
//   model_RGB_asympt_aj_CteWidth_HarveyLike_v4   tamcmc/sources/models.cpp:4334-4682   (id 27)
// i.e. everything those functions do BEFORE their optimum_lorentzian_calc_aj loops: Appourchaux / constant l=0 widths, the
// asymptotic mixed-mode solver for the l=1 frequencies (ARMM), the bias spline, the zeta function, heights, widths and
// two-zone splittings of the mixed modes.  The result is one MODE TABLE row (include/tamcmc_gpu.h): exactly the arguments the
// reference passes to optimum_lorentzian_calc_aj (models.cpp:4931-5006); windows, profiles, noise and likelihood run on the GPU.
// Part of libtamcmc_gpu.so; plain host C++ (no CUDA calls), compiled with -ffp-contract=off.
//
// Reference pieces restated here (new code; same operations in the same order and in the same floating-point types, because
// a mixed mode can be narrower than 1e-3 microHz: its frequency has to agree to ~1e-15 relative for 1e-10 on the spectrum):
//   sign_change, pnu_fct, gnu_fct, asympt_nu_p, asympt_nu_p_from_l0_Xd, asympt_nu_g, solver_mm,
//   solve_mm_asymptotic_O2p, solve_mm_asymptotic_O2from_l0        external/ARMM/solver_mm.cpp:71-740
//   Frstder_adaptive_reggrid                                       external/ARMM/derivatives_handler.cpp:425-444
//   where_in_range, where_dbl                                      external/ARMM/string_handler.cpp:135-237
//   ksi_fct1, ksi_fct2 ("precise"), gamma_l_fct2, h_l_rgb, dnu_rot_2zones   external/ARMM/bump_DP.cpp:46-240, 531-537
//   tk::spline (cspline / cspline_hermite, second-derivative boundaries)    external/spline/src/spline.h:219-405, 480-503, 685-761
//   linfit, lin_interpol, eta0_fct                                 tamcmc/sources/linfit.cpp:16-33, interpol.cpp:13-43, models.cpp:6065-6084
// Eigen is not used: a `VectorXd op long double` expression of the reference converts the scalar to double and applies the
// operation element by element, which is what the loops below do (the casts are where the reference's conversions are).
// The reference runs the (p mode, g mode) pairs and the zeta sums under OpenMP with critical sections (solver_mm.cpp:561,
// bump_DP.cpp:137,153): its own summation order depends on the thread schedule.  Here pairs are independent and their
// solutions are sorted afterwards, and the zeta sums are taken per frequency in (np, ng) order -- the single-thread order of
// the reference -- so results do not depend on the thread count.
#include "../../include/tamcmc_gpu.h"
#include "host_math.hpp"
#include "host_rgb.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

typedef long double ld;
typedef std::vector<double> vec;

// Eigen::VectorXd::LinSpaced(n, lo, hi): lo + i * (hi - lo) / (n - 1), last element = hi
vec linspaced(long n, double lo, double hi)
{
    vec r((size_t)(n > 0 ? n : 0));
    if (n <= 0) return r;
    if (n == 1) { r[0] = hi; return r; }
    const double step = (hi - lo) / (double)(n - 1);
    for (long i = 0; i < n; i++) r[(size_t)i] = lo + (double)i * step;
    r[(size_t)n - 1] = hi;
    return r;
}

// linfit (tamcmc/sources/linfit.cpp:16-33): out[0] slope, out[1] intercept
void linfit(const vec& x, const vec& y, double out[2])
{
    double sx = 0, sy = 0;
    for (double v : x) sx += v;
    for (double v : y) sy += v;
    const double n = (double)x.size();
    const double mean_x = sx / n;
    double sty = 0, stt = 0;
    for (size_t i = 0; i < x.size(); i++) { const double t = x[i] - mean_x; sty += t * y[i]; }
    for (size_t i = 0; i < x.size(); i++) { const double t = x[i] - mean_x; stt += t * t; }
    out[0] = sty / stt;
    out[1] = (sy - sx * out[0]) / n;
}

double vmin(const vec& v) { double m = v[0]; for (double a : v) if (a < m) m = a; return m; }
double vmax(const vec& v) { double m = v[0]; for (double a : v) if (a > m) m = a; return m; }

// ---------------------------------------------------------------------------------------------- solver_mm.cpp
// sign_change (solver_mm.cpp:71-106): positions i where x[i] -> x[i+1] crosses (or touches) zero
void sign_change(const vec& x, std::vector<long>& pos)
{
    pos.clear();
    for (size_t i = 0; i + 1 < x.size(); i++) {
        if ((x[i + 1] >= 0 && x[i] < 0) || (x[i + 1] > 0 && x[i] <= 0)) pos.push_back((long)i);       // - to +
        if ((x[i + 1] <= 0 && x[i] > 0) || (x[i + 1] <= 0 && x[i] >= 0)) pos.push_back((long)i);      // + to -
    }
}

// p(nu) - g(nu): pnu_fct (solver_mm.cpp:114-127) minus gnu_fct (solver_mm.cpp:150-161) as their VectorXd versions compute one
// element (double arithmetic, the long double scalars converted first)
struct PminusG {
    double nu_p_d, inv_g, pi_d, DPl_d, q_d, Dnu_d;
    PminusG(ld nu_p, ld nu_g, ld Dnu_p, ld DPl, ld q)
        : nu_p_d((double)nu_p), inv_g(1.0 / (double)nu_g), pi_d((double)3.141592653589793238L), DPl_d((double)DPl), q_d((double)q), Dnu_d((double)Dnu_p) {}
    double operator()(double nu) const
    {
        const double pnu = nu - nu_p_d;
        const double X = ((pi_d * (1.0 / nu - inv_g)) * 1e6) / DPl_d;
        const double t = q_d * std::tan(X);
        const double gnu = (Dnu_d * std::atan(t)) / pi_d;
        return pnu - gnu;
    }
    // X / pi as a function of nu (decreasing), and its inverse: the tangent has its poles where this is m + 1/2, m integer.  Between
    // two poles g falls from +Dnu/2 to -Dnu/2, so p - g rises with a slope > 1: it changes sign at most once there, from - to +;
    // at a pole it jumps from + to -.  (Only used to LOCATE poles to within a grid step; every sign is taken from operator().)
    double u_of(double nu) const { return (1.0 / nu - inv_g) * 1e6 / DPl_d; }
    double nu_of_u(double u) const { return 1.0 / (u * DPl_d / 1e6 + inv_g); }
    bool usable() const { return DPl_d > 0 && q_d > 0 && Dnu_d > 0 && std::isfinite(inv_g) && inv_g > 0; }
};

// Does a pole of the tangent lie inside [a, b] (with a safety margin of `pad` on both sides)?
bool pole_inside(const PminusG& F, double a, double b, double pad)
{
    const double lo = a - pad, hi = b + pad;
    if (!(lo > 0)) return true;
    const double u_hi = F.u_of(lo), u_lo = F.u_of(hi);          // u decreases with nu
    return std::floor(u_hi - 0.5) >= std::ceil(u_lo - 0.5);
}

// lin_interpol(f(nu_local), nu_local, 0) (tamcmc/sources/interpol.cpp:13-43) with f evaluated ON DEMAND: the function reads
// both end values, then either walks from the left until it brackets zero (an increasing crossing) or -- when f runs from + to
// -, which is what the jump of g at a pole of the tangent looks like -- only the first and the last two values.  Same
// comparisons and arithmetic as lin_interpol on the full vector; a fraction of the tan / atan calls.
double interp_zero_lazy(const vec& y /*nu_local*/, const PminusG& F, vec& val, std::vector<unsigned char>& have, bool monotone)
{
    const int Nx = (int)y.size();
    val.assign((size_t)Nx, 0.0); have.assign((size_t)Nx, 0);
    auto X = [&](int i) { if (!have[(size_t)i]) { val[(size_t)i] = F(y[(size_t)i]); have[(size_t)i] = 1; } return val[(size_t)i]; };
    const double x_int = 0.0;
    int i = 0;
    double a = 0, b = 0;
    if (x_int >= X(0) && x_int <= X(Nx - 1)) {
        // The reference walks from the left to the first i with X(i) <= 0 <= X(i+1).  Where p - g is monotone on the whole local
        // range (no pole of the tangent inside it) and no value is exactly zero, that i is the one sign change: found by bisection,
        // with the same values in the same comparisons at the end.  Anything else (a pole in range, an exact zero): the walk.
        if (monotone && X(0) < 0.0 && X(Nx - 1) > 0.0 && Nx >= 3) {
            int lo = 0, hi = Nx - 1;
            bool exact_zero = false;
            while (hi - lo > 1) {
                const int mid = (lo + hi) / 2;
                const double v = X(mid);
                if (v == 0.0) { exact_zero = true; break; }
                if (v < 0.0) lo = mid; else hi = mid;
            }
            if (!exact_zero) i = std::min(lo, Nx - 2);
        }
        while ((x_int < X(i) || x_int > X(i + 1)) && i < Nx - 2) i = i + 1;
        a = (y[(size_t)i + 1] - y[(size_t)i]) / (X(i + 1) - X(i));
        b = y[(size_t)i] - a * X(i);
    }
    if (x_int < X(0)) { a = (y[1] - y[0]) / (X(1) - X(0)); b = y[0] - a * X(0); }
    if (x_int > X(Nx - 1)) { a = (y[(size_t)Nx - 1] - y[(size_t)Nx - 2]) / (X(Nx - 1) - X(Nx - 2)); b = y[(size_t)Nx - 2] - a * X(Nx - 2); }
    return a * x_int + b;
}

// gnu_fct for one frequency (solver_mm.cpp:172-180), long double throughout
ld gnu_scalar(ld nu, ld nu_g, ld Dnu_p, ld DPl, ld q)
{
    const ld pi = 3.141592653589793238L;
    const ld X = pi * (1. / nu - 1. / nu_g) * 1e6 / DPl;
    return Dnu_p * atanl(q * tanl(X)) / pi;
}

// TAMCMC_ARMM_EXACT_SCAN=1 (read once): evaluate p - g on every grid point like the reference does, instead of locating the sign
// changes by bisection between the poles of the tangent (same indices, same solutions: tests/test_rgb_expander.py runs both).
const bool g_armm_fast = [] { const char* e = std::getenv("TAMCMC_ARMM_EXACT_SCAN"); return !(e && e[0] == '1'); }();

// The indices sign_change() returns for f = p - g on the grid `nu`, without evaluating f everywhere: the poles of the tangent cut
// the band into stretches where f rises monotonically (one - to + change at most, found by bisection on the grid index); around
// each pole the four nearest grid values are taken and compared pair by pair exactly like sign_change does (the + to - jump).
// Returns false -- and the caller falls back to the full scan -- when anything is not as assumed: an exact zero, a value that is
// not finite, a stretch whose ends are not ordered like a rising function, too many poles for the grid.
bool band_sign_changes(const vec& nu, const PminusG& F, vec& f, std::vector<long>& idx)
{
    idx.clear();
    const long n = (long)nu.size();
    if (!g_armm_fast || !F.usable() || n < 8) return false;
    const double gstep = (nu[(size_t)n - 1] - nu[0]) / (double)(n - 1);
    if (!(gstep > 0) || !(nu[0] > 0)) return false;
    // boundaries: the band's ends and the grid neighbours of every pole inside the band
    const double u_hi = F.u_of(nu[0]), u_lo = F.u_of(nu[(size_t)n - 1]);
    const double m_hi = std::floor(u_hi - 0.5), m_lo = std::ceil(u_lo - 0.5);
    if (!(std::isfinite(m_hi) && std::isfinite(m_lo)) || m_hi - m_lo > (double)n / 6.0) return false;      // poles closer than ~6 grid steps: scan
    std::vector<long> B;
    B.push_back(0); B.push_back(n - 1);
    for (double m = m_hi; m >= m_lo; m -= 1.0) {
        const double nup = F.nu_of_u(m + 0.5);
        const long ip = (long)std::floor((nup - nu[0]) / gstep);
        for (long k = ip - 1; k <= ip + 2; k++) if (k >= 0 && k <= n - 1) B.push_back(k);
    }
    std::sort(B.begin(), B.end());
    B.erase(std::unique(B.begin(), B.end()), B.end());
    std::vector<unsigned char> have((size_t)n, 0);
    bool ok = true;
    auto X = [&](long i) { if (!have[(size_t)i]) { f[(size_t)i] = F(nu[(size_t)i]); have[(size_t)i] = 1; if (!(f[(size_t)i] != 0.0) || !std::isfinite(f[(size_t)i])) ok = false; } return f[(size_t)i]; };
    for (size_t k = 0; k + 1 < B.size() && ok; k++) {
        const long a = B[k], b = B[k + 1];
        const double fa = X(a), fb = X(b);
        if (!ok) break;
        if (b == a + 1) {
            // neighbours: the two tests of sign_change (solver_mm.cpp:71-106); no value is zero here, so at most one fires
            if (fb > 0 && fa < 0) idx.push_back(a);
            if (fb < 0 && fa > 0) idx.push_back(a);
        } else {
            // a stretch without a pole: rising
            if (fa > 0 && fb < 0) { ok = false; break; }                    // not what a rising function does: scan instead
            if (fa < 0 && fb > 0) {
                long lo = a, hi = b;
                while (hi - lo > 1 && ok) { const long mid = lo + (hi - lo) / 2; if (X(mid) < 0) lo = mid; else hi = mid; }
                if (ok) idx.push_back(lo);
            }
        }
    }
    if (!ok) { idx.clear(); return false; }
    return true;
}

// solver_mm (solver_mm.cpp:326-449): intersections of p(nu) = nu - nu_p and g(nu) for ONE (p mode, g mode) pair.
// The reference evaluates p - g on the whole grid [numin, numax] (3.5 large separations at the spectrum's resolution) and keeps
// the sign changes.  |g(nu)| = |Dnu atan(.)/pi| <= Dnu/2, so outside |nu - nu_p| <= Dnu/2 the difference has the sign of nu - nu_p
// and cannot change sign: only the grid points of that band (plus a margin of a few points) are evaluated -- the same grid
// values, hence the same sign-change indices and the same solutions, for less than a third of the tan / atan calls.
// The coarse grid of one p mode and the band of it that can hold a sign change (see solver_mm below)
struct BandGrid {
    long n = 0, i_lo = 0, i_hi = -1;
    double lo = 0, hi = 0, gstep = 0;
    bool valid = false;
    double grid(long i) const { return (i == n - 1) ? hi : lo + (double)i * gstep; }      // Eigen::VectorXd::LinSpaced(n, lo, hi)[i]
};
BandGrid band_of(ld nu_p, ld Dnu_p, ld numin, ld numax, ld resol)
{
    BandGrid B;
    B.n = (numin >= 0) ? (long)((numax - numin) / resol) : (long)((numax) / resol);
    B.lo = (numin >= 0) ? (double)numin : 0.0; B.hi = (double)numax;
    if (B.n < 2) return B;
    B.gstep = (B.hi - B.lo) / (double)(B.n - 1);
    B.i_lo = 0; B.i_hi = B.n - 1;
    const double half = 0.5 * std::fabs((double)Dnu_p) * (1.0 + 1e-9) + 4.0 * std::fabs(B.gstep);
    const double a = ((double)nu_p - half - B.lo) / B.gstep, b = ((double)nu_p + half - B.lo) / B.gstep;
    if (std::isfinite(a) && std::isfinite(b) && B.gstep > 0) {
        B.i_lo = std::max(0L, (long)std::floor(a) - 2);
        B.i_hi = std::min(B.n - 1, (long)std::ceil(b) + 2);
    }
    B.valid = !(B.i_hi < B.i_lo + 1);                                                    // else the band lies outside the grid: no sign change
    return B;
}
// the local grid the reference refines a sign change at coarse frequency nu_idx on (solver_mm.cpp:402-410): long double arithmetic
inline long local_grid(double nu_idx, ld resol, ld factor, double& range_min_d, double& range_max_d)
{
    const ld range_min = nu_idx - 2 * resol, range_max = nu_idx + 2 * resol;
    range_min_d = (double)range_min; range_max_d = (double)range_max;
    return (long)((range_max - range_min) / (resol * factor));
}

// solver_mm (solver_mm.cpp:326-449): intersections of p(nu) = nu - nu_p and g(nu) for ONE (p mode, g mode) pair.
// The reference evaluates p - g on the whole grid [numin, numax] (3.5 large separations at the spectrum's resolution) and keeps
// the sign changes.  |g(nu)| = |Dnu atan(.)/pi| <= Dnu/2, so outside |nu - nu_p| <= Dnu/2 the difference has the sign of nu - nu_p
// and cannot change sign: only the grid points of that band (plus a margin of a few points) are evaluated -- the same grid
// values, hence the same sign-change indices and the same solutions, for less than a third of the tan / atan calls.
void solver_mm(ld nu_p, ld nu_g, ld Dnu_p, ld DPl, ld q, ld numin, ld numax, ld resol, ld factor, vec& nu_m)
{
    nu_m.clear();
    if (!(nu_g >= numin && nu_g <= numax)) return;
    const BandGrid G = band_of(nu_p, Dnu_p, numin, numax, resol);
    if (!G.valid) return;
    const long i_lo = G.i_lo, i_hi = G.i_hi;
    vec nu((size_t)(i_hi - i_lo + 1));
    for (long i = i_lo; i <= i_hi; i++) nu[(size_t)(i - i_lo)] = G.grid(i);
    vec f(nu.size()), nu_local, f_local;
    std::vector<unsigned char> have;
    std::vector<long> idx;
    const PminusG F(nu_p, nu_g, Dnu_p, DPl, q);
    if (!band_sign_changes(nu, F, f, idx)) {             // the reference's own scan: every grid value, then sign_change
        for (size_t i = 0; i < nu.size(); i++) f[i] = F(nu[i]);
        sign_change(f, idx);
    }
    for (size_t k = 0; k < idx.size(); k++) {
        double rmin, rmax;
        const long nloc = local_grid(nu[(size_t)idx[k]], resol, factor, rmin, rmax);
        nu_local = linspaced(nloc, rmin, rmax);
        if (nu_local.size() < 2) continue;
        const bool monotone = g_armm_fast && F.usable() && !pole_inside(F, nu_local.front(), nu_local.back(), 4.0 * (double)(resol * factor));
        const ld nu_m_proposed = interp_zero_lazy(nu_local, F, f_local, have, monotone);
        const ld ysol_gnu = gnu_scalar(nu_m_proposed, nu_g, Dnu_p, DPl, q);
        const ld ysol_pnu = nu_m_proposed - nu_p;
        const ld ratio = ysol_gnu / ysol_pnu;
        // only intersections that really satisfy p = g (to 0.1 %) are kept: the arctan branch cuts produce spurious sign changes
        if ((ratio >= 0.999) && (ratio <= 1.001)) nu_m.push_back((double)nu_m_proposed);
    }
}

struct Eigensols { vec nu_m, nu_p, nu_g, dnup, dPg; bool ok = false; };

// Frstder_adaptive_reggrid(y) (derivatives_handler.cpp:425-444): one-sided differences at the ends, centred inside
vec first_derivative(const vec& y)
{
    const size_t n = y.size();
    vec d(n, 0.0);
    if (n < 2) return d;
    d[0] = y[1] - y[0];
    d[n - 1] = y[n - 1] - y[n - 2];
    for (size_t i = 0; i + 2 < n; i++) d[i + 1] = (y[i + 2] - y[i]) / 2.;
    return d;
}

// keep what is inside [keep_min, keep_max], sort, std::unique with tolerance (solver_mm.cpp:575-590, 722-740)
void sort_unique(vec& all, ld resol, ld keep_min, ld keep_max, vec& nu_m_all)
{
    vec kept;
    for (double s : all)
        if ((ld)s >= keep_min && (ld)s <= keep_max) kept.push_back(s);
    std::sort(kept.begin(), kept.end());
    const double tol = (double)(2 * resol);
    nu_m_all.clear();
    for (double s : kept)                                           // std::unique: compare with the last element KEPT
        if (nu_m_all.empty() || !(std::abs(nu_m_all.back() - s) <= tol)) nu_m_all.push_back(s);
}

// the arguments of solve_pairs, kept when the pair loop is to run somewhere else (the device solver, rgb_device.cu)
struct PairSetup { vec dnu_local; std::vector<ld> lo, hi; ld DPl = 0, q = 0, resol = 0; double fact = 0; ld keep_min = 0, keep_max = 0; bool set = false; };

// the pair loop shared by the two entry points + sort + std::unique with tolerance (solver_mm.cpp:558-590, 706-740)
// dnu_local[np]: the local large separation handed to solver_mm; centre +- zone[np] is the search range
void solve_pairs(const vec& nu_p_all, const vec& nu_g_all, const vec& dnu_local, const std::vector<ld>& lo, const std::vector<ld>& hi, ld DPl, ld q,
                 ld resol, double fact, ld keep_min, ld keep_max, vec& nu_m_all)
{
    const long Np = (long)nu_p_all.size(), Ng = (long)nu_g_all.size();
    std::vector<vec> found((size_t)(Np * Ng));
#ifdef _OPENMP
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
#endif
    for (long np = 0; np < Np; np++)
        for (long ng = 0; ng < Ng; ng++)
            solver_mm(nu_p_all[(size_t)np], nu_g_all[(size_t)ng], dnu_local[(size_t)np], DPl, q, lo[(size_t)np], hi[(size_t)np], resol, fact,
                      found[(size_t)(np * Ng + ng)]);
    vec all;
    for (const vec& v : found)
        for (double s : v) all.push_back(s);
    sort_unique(all, resol, keep_min, keep_max, nu_m_all);
}

// ng range and search parameters common to both entry points (solver_mm.cpp:497-533, 649-683)
bool g_mode_setup(ld fmin, ld fmax, ld DPl, ld alpha, int& ng_min, int& ng_max, double& fact)
{
    ng_min = (int)floorl(1e6 / (fmax * DPl) - alpha);
    ng_max = (int)ceill(1e6 / (fmin * DPl) - alpha);
    if (ng_min <= 0 && ng_max < 1) return false;                    // "You requested an impossible star": the reference returns nothing
    if (ng_min <= 0 && ng_max >= 1) ng_min = 1;
    fact = 0.04;
    if (fmin <= 150) fact = 0.01;
    if (fmin <= 50) fact = 0.005;
    return true;
}

// solve_mm_asymptotic_O2from_l0 (solver_mm.cpp:624-746), sigma_p = 0
void solve_from_l0(const vec& nu_l0_in, int el, ld delta0l, ld DPl, ld alpha, ld q, ld resol, ld freq_min, ld freq_max, Eigensols& S, PairSetup* defer = nullptr)
{
    S = Eigensols();
    const size_t n0 = nu_l0_in.size();
    double fit[2];
    linfit(linspaced((long)n0, 0.0, (double)(n0 - 1)), nu_l0_in, fit);
    const double Dnu_p = fit[0];
    double fmin = vmin(nu_l0_in) - Dnu_p, fmax = vmax(nu_l0_in) + Dnu_p;
    if (fmin < 0) fmin = 0;
    int ng_min, ng_max; double fact;
    if (!g_mode_setup(fmin, fmax, DPl, alpha, ng_min, ng_max, fact)) return;
    const double Coeff = (ng_max - ng_min < 6) ? 20 : 1.75;
    // asympt_nu_p_from_l0_Xd (solver_mm.cpp:263-304): the l=0 comb extended by three orders on each side, shifted to degree el
    {
        const ld Dnu = Dnu_p;
        vec l0_long(n0 + 6);
        l0_long[0] = (double)(vmin(nu_l0_in) - 3 * Dnu); l0_long[1] = (double)(vmin(nu_l0_in) - 2 * Dnu); l0_long[2] = (double)(vmin(nu_l0_in) - Dnu);
        for (size_t k = 0; k < n0; k++) l0_long[k + 3] = nu_l0_in[k];
        l0_long[n0 + 3] = (double)(vmax(nu_l0_in) + Dnu); l0_long[n0 + 4] = (double)(vmax(nu_l0_in) + 2 * Dnu); l0_long[n0 + 5] = (double)(vmax(nu_l0_in) + 3 * Dnu);
        const double shift = (double)(el / 2. * Dnu + delta0l);
        const double lo = (double)(ld)fmin, hi = (double)(ld)fmax;               // (fmin == -1 / fmax == -1 defaults do not occur: both are >= 0 here)
        for (double v : l0_long) { const double s = v + shift; if (s >= lo && s <= hi) S.nu_p.push_back(s); }
        if (S.nu_p.empty()) return;                                             // the reference exits: "No frequency found in the specified range"
    }
    for (int ng = ng_min; ng < ng_max; ng++) S.nu_g.push_back((double)(1e6 / ((ng + alpha) * DPl + 0)));     // asympt_nu_g, solver_mm.cpp:314-318
    S.dnup = first_derivative(S.nu_p);
    S.dPg.assign(S.nu_g.size(), (double)DPl);
    std::vector<ld> lo(S.nu_p.size()), hi(S.nu_p.size());
    for (size_t np = 0; np < S.nu_p.size(); np++) { lo[np] = S.nu_p[np] - Coeff * Dnu_p; hi[np] = S.nu_p[np] + Coeff * Dnu_p; }     // double arithmetic
    if (defer) { *defer = PairSetup{S.dnup, lo, hi, DPl, q, resol, fact, freq_min, freq_max, true}; S.ok = true; return; }
    solve_pairs(S.nu_p, S.nu_g, S.dnup, lo, hi, DPl, q, resol, fact, freq_min, freq_max, S.nu_m);
    S.ok = true;
}

// solve_mm_asymptotic_O2p (solver_mm.cpp:470-604), sigma_p = 0
int solve_O2p(ld Dnu_p, ld epsilon, int el, ld delta0l, ld alpha_p, ld nmax, ld DPl, ld alpha, ld q, ld fmin, ld fmax, ld resol, Eigensols& S, PairSetup* defer = nullptr)
{
    S = Eigensols();
    int np_min = (int)floorl(fmin / Dnu_p - epsilon - el / 2 - delta0l);         // el / 2: integer division like the reference
    int np_max = (int)ceill(fmax / Dnu_p - epsilon - el / 2 - delta0l);
    np_min = (int)floorl(np_min - alpha_p * powl(np_min - nmax, 2) / 2.);
    np_max = (int)ceill(np_max + alpha_p * powl(np_max - nmax, 2) / 2.);
    int ng_min, ng_max; double fact;
    if (!g_mode_setup(fmin, fmax, DPl, alpha, ng_min, ng_max, fact)) return TAMCMC_OK;
    const double Coeff = (ng_max - ng_min < 6) ? (double)np_max : 1.75;
    if (np_min <= 0) np_min = 1;
    for (int np = np_min; np < np_max; np++) {
        const ld nu_p = (np + epsilon + el / 2. + delta0l + alpha_p * powl(np - nmax, 2) / 2) * Dnu_p;     // asympt_nu_p, solver_mm.cpp:201-215
        if (nu_p < 0.0) return TAMCMC_ERR_NONFINITE;                                                       // the reference exits
        S.nu_p.push_back((double)(nu_p + 0));
    }
    for (int ng = ng_min; ng < ng_max; ng++) S.nu_g.push_back((double)(1e6 / ((ng + alpha) * DPl + 0)));
    if (S.nu_p.empty()) { S.ok = true; return TAMCMC_OK; }
    S.dnup = first_derivative(S.nu_p);
    S.dPg.assign(S.nu_g.size(), (double)DPl);
    vec dnu_local(S.nu_p.size());
    std::vector<ld> lo(S.nu_p.size()), hi(S.nu_p.size());
    for (size_t np = 0; np < S.nu_p.size(); np++) {
        dnu_local[np] = (double)(Dnu_p * (1.0 + alpha_p * (np + np_min - nmax)));
        lo[np] = S.nu_p[np] - Coeff * Dnu_p; hi[np] = S.nu_p[np] + Coeff * Dnu_p;                           // long double arithmetic (Dnu_p is)
    }
    if (defer) { *defer = PairSetup{dnu_local, lo, hi, DPl, q, resol, fact, fmin, fmax, true}; S.ok = true; return TAMCMC_OK; }
    solve_pairs(S.nu_p, S.nu_g, dnu_local, lo, hi, DPl, q, resol, fact, fmin, fmax, S.nu_m);
    S.ok = true;
    return TAMCMC_OK;
}

// ---------------------------------------------------------------------------------------------- bump_DP.cpp
// Sum over (np, ng) of ksi_fct1 (bump_DP.cpp:46-64) at every frequency of `nu`, in the reference's single-thread order: for each
// np a local sum over ng, then added to the total (bump_DP.cpp:137-151).  One term is
//     1 / (1 + front * cos^2(up) / cos^2(down)),   up = (pi 1e6 (1/nu - 1/nu_g)) / DPl,  down = (pi (nu - nu_p)) / Dnu_p,
//     front = ((1e-6 nu^2) DPl) / (q Dnu_p):
// `up` depends on (nu, ng) only and `down` on (nu, np) only, so the Lp + Lg cosines of a frequency are taken once instead of
// 2 Lp Lg times -- the same arguments give the same cosines, every other operation is unchanged.
// W frequencies at a time: the cosines of a frequency stay scalar libm calls, the Lp x Lg divisions run in SIMD lanes, one lane per
// frequency -- every lane performs exactly the scalar sequence of IEEE operations in the scalar order (per np a sum over ng, then
// added to the total), so the results do not depend on the vector width.  Compiled for AVX-512 / AVX2 / baseline x86-64 with
// run-time dispatch (the library must load on any host).
constexpr int KSI_W = 8;
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target_clones("avx512f", "avx2", "default")))
#endif
void ksi_block(size_t Lp, size_t Lg, const double* cu2 /*[Lg][W]*/, const double* nd /*[Lg][W]*/, const double* cd2 /*[Lp][W]*/,
               const double* qD /*[Lp]*/, double* tot /*[W]*/)
{
    double t[KSI_W];
    for (int w = 0; w < KSI_W; w++) t[w] = 0.0;
    for (size_t p = 0; p < Lp; p++) {
        double loc[KSI_W];
        for (int w = 0; w < KSI_W; w++) loc[w] = 0.0;
        const double qd = qD[p];
        const double* cd = cd2 + p * KSI_W;
        for (size_t g = 0; g < Lg; g++) {
            const double* cu = cu2 + g * KSI_W;
            const double* n_ = nd + g * KSI_W;
#pragma omp simd
            for (int w = 0; w < KSI_W; w++) loc[w] += 1.0 / (1.0 + (n_[w] / qd) * (cu[w] / cd[w]));
        }
        for (int w = 0; w < KSI_W; w++) t[w] += loc[w];
    }
    for (int w = 0; w < KSI_W; w++) tot[w] = t[w];
}

// the per-(p mode), per-(g mode) constants of the zeta sums, and the sums of ONE block of KSI_W frequencies
struct KsiConst {
    std::vector<double> inv_g, qD;
    double c_up = 0, pi_d = 0;
    KsiConst(const vec& nu_g, const vec& Dnu_p, ld q)
    {
        const ld pi = M_PI;
        c_up = (double)(pi * 1e6); pi_d = (double)pi;
        inv_g.resize(nu_g.size()); qD.resize(Dnu_p.size());
        for (size_t g = 0; g < nu_g.size(); g++) inv_g[g] = (double)(1. / (ld)nu_g[g]);
        for (size_t p = 0; p < Dnu_p.size(); p++) qD[p] = (double)(q * (ld)Dnu_p[p]);
    }
};
struct KsiScratch { std::vector<double> cu2, cd2, nd; };
void ksi_sum_block(long b, const vec& nu, const vec& nu_p, const vec& Dnu_p, const vec& DPl, const KsiConst& K, KsiScratch& W, vec& out)
{
    const size_t Lp = nu_p.size(), Lg = K.inv_g.size();
    const long N = (long)nu.size();
    W.cu2.resize(Lg * KSI_W); W.cd2.resize(Lp * KSI_W); W.nd.resize(Lg * KSI_W);
    double tot[KSI_W];
    for (int w = 0; w < KSI_W; w++) {
        const long i = b * KSI_W + w;
        if (i > N - 1 && w > 0) {                                // the last block repeats its last frequency in the spare lanes (copied, not recomputed)
            for (size_t g = 0; g < Lg; g++) { W.cu2[g * KSI_W + w] = W.cu2[g * KSI_W + w - 1]; W.nd[g * KSI_W + w] = W.nd[g * KSI_W + w - 1]; }
            for (size_t p = 0; p < Lp; p++) W.cd2[p * KSI_W + w] = W.cd2[p * KSI_W + w - 1];
            continue;
        }
        const double v = nu[(size_t)std::min(i, N - 1)];
        const double inv = 1.0 / v, sq = 1e-6 * (v * v);
        for (size_t g = 0; g < Lg; g++) { const double c = std::cos((K.c_up * (inv - K.inv_g[g])) / DPl[g]); W.cu2[g * KSI_W + w] = c * c; W.nd[g * KSI_W + w] = sq * DPl[g]; }
        for (size_t p = 0; p < Lp; p++) { const double c = std::cos((K.pi_d * (v - nu_p[p])) / Dnu_p[p]); W.cd2[p * KSI_W + w] = c * c; }
    }
    ksi_block(Lp, Lg, W.cu2.data(), W.nd.data(), W.cd2.data(), K.qD.data(), tot);
    for (int w = 0; w < KSI_W; w++) { const long i = b * KSI_W + w; if (i < N) out[(size_t)i] = tot[w]; }
}

void ksi_sum(const vec& nu, const vec& nu_p, const vec& nu_g, const vec& Dnu_p, const vec& DPl, ld q, vec& out)
{
    const KsiConst K(nu_g, Dnu_p, q);
    out.assign(nu.size(), 0.0);
    const long N = (long)nu.size();
    const long NB = (N + KSI_W - 1) / KSI_W;
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        KsiScratch W;
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (long b = 0; b < NB; b++) ksi_sum_block(b, nu, nu_p, Dnu_p, DPl, K, W, out);
    }
}

// ksi_fct2(..., "precise") = ksi_fct2_precise (bump_DP.cpp:126-177): normalised by the maximum over a 4-year-resolution grid
// the high-resolution grid of ksi_fct2_precise (bump_DP.cpp:126-135): [fmin, fmax] over the p and g modes at 4-year resolution
bool ksi_highres_grid(const vec& nu_p, const vec& nu_g, double& fmin_d, double& fmax_d, int& Ndata)
{
    if (nu_p.empty() || nu_g.empty()) return false;
    const ld resol = 1e6 / (4 * 365. * 86400.);
    const ld fmin = (vmin(nu_p) >= vmin(nu_g)) ? vmin(nu_g) : vmin(nu_p);
    const ld fmax = (vmax(nu_p) >= vmax(nu_g)) ? vmax(nu_p) : vmax(nu_g);
    Ndata = (int)((fmax - fmin) / resol);
    fmin_d = (double)fmin; fmax_d = (double)fmax;
    return Ndata >= 1;
}

// norm_in >= 0: the maximum over the high-resolution grid computed elsewhere (the device, rgb_device.cu)
// have_sums: ksi_pg already holds the zeta sums at `nu` (computed block by block elsewhere)
bool ksi_fct2_precise(const vec& nu, const vec& nu_p, const vec& nu_g, const vec& Dnu_p, const vec& DPl, ld q, vec& ksi_pg, double norm_in = -1.0,
                      bool have_sums = false)
{
    double fmin, fmax; int Ndata;
    if (!ksi_highres_grid(nu_p, nu_g, fmin, fmax, Ndata)) return false;
    if (!have_sums) ksi_sum(nu, nu_p, nu_g, Dnu_p, DPl, q, ksi_pg);
    ld norm_coef = norm_in;
    if (!(norm_in >= 0)) {
        const vec nu_highres = linspaced(Ndata, fmin, fmax);
        vec ksi_highres;
        ksi_sum(nu_highres, nu_p, nu_g, Dnu_p, DPl, q, ksi_highres);
        norm_coef = vmax(ksi_highres);
    }
    for (double& v : ksi_pg) { v = v / (double)norm_coef; if (v > 1) v = 1; }
    return true;
}

// ---------------------------------------------------------------------------------------------- tk::spline
// cubic spline through (x, y) with f'' = 0 at both ends: type 1 = cspline (C2), type 2 = cspline_hermite (C1, 3-point slopes)
struct Spline {
    vec x, y, b, c, d;
    double c0 = 0.0;
    void set_coeffs_from_b()                                                       // spline.h:219-240
    {
        const size_t n = b.size();
        c.resize(n); d.resize(n);
        for (size_t i = 0; i + 1 < n; i++) {
            const double h = x[i + 1] - x[i];
            c[i] = (3.0 * (y[i + 1] - y[i]) / h - (2.0 * b[i] + b[i + 1])) / h;
            d[i] = ((b[i + 1] - b[i]) / (3.0 * h) - 2.0 / 3.0 * c[i]) / h;
        }
        c0 = c[0];
    }
    bool set_points(const vec& xs, const vec& ys, int type)                        // spline.h:242-412
    {
        x = xs; y = ys;
        const int n = (int)x.size();
        if (n < 3) return false;
        for (int i = 0; i + 1 < n; i++) if (!(x[(size_t)i] < x[(size_t)i + 1])) return false;
        if (type == 1) {
            // tridiagonal system for c[] (band_matrix with one lower and one upper diagonal), solved like band_matrix::lu_solve
            // (spline.h:685-761): rows scaled to a unit diagonal, Gauss elimination without pivoting, two triangular solves
            vec lo((size_t)n, 0.0), di((size_t)n, 0.0), up((size_t)n, 0.0), rhs((size_t)n, 0.0), sd((size_t)n, 0.0);
            for (int i = 1; i < n - 1; i++) {
                lo[(size_t)i] = 1.0 / 3.0 * (x[(size_t)i] - x[(size_t)i - 1]);
                di[(size_t)i] = 2.0 / 3.0 * (x[(size_t)i + 1] - x[(size_t)i - 1]);
                up[(size_t)i] = 1.0 / 3.0 * (x[(size_t)i + 1] - x[(size_t)i]);
                rhs[(size_t)i] = (y[(size_t)i + 1] - y[(size_t)i]) / (x[(size_t)i + 1] - x[(size_t)i]) - (y[(size_t)i] - y[(size_t)i - 1]) / (x[(size_t)i] - x[(size_t)i - 1]);
            }
            di[0] = 2.0; up[0] = 0.0; rhs[0] = 0.0;                                 // 2 c[0] = f'' = 0
            di[(size_t)n - 1] = 2.0; lo[(size_t)n - 1] = 0.0; rhs[(size_t)n - 1] = 0.0;
            for (int i = 0; i < n; i++) {
                sd[(size_t)i] = 1.0 / di[(size_t)i];
                if (i >= 1) lo[(size_t)i] *= sd[(size_t)i];
                if (i < n - 1) up[(size_t)i] *= sd[(size_t)i];
                di[(size_t)i] = 1.0;
            }
            for (int k = 0; k + 1 < n; k++) {
                const double f = -lo[(size_t)k + 1] / di[(size_t)k];
                lo[(size_t)k + 1] = -f;
                di[(size_t)k + 1] = di[(size_t)k + 1] + f * up[(size_t)k];
            }
            vec yy((size_t)n);
            for (int i = 0; i < n; i++) { double sum = 0; if (i >= 1) sum += lo[(size_t)i] * yy[(size_t)i - 1]; yy[(size_t)i] = (rhs[(size_t)i] * sd[(size_t)i]) - sum; }
            c.assign((size_t)n, 0.0);
            for (int i = n - 1; i >= 0; i--) { double sum = 0; if (i < n - 1) sum += up[(size_t)i] * c[(size_t)i + 1]; c[(size_t)i] = (yy[(size_t)i] - sum) / di[(size_t)i]; }
            b.assign((size_t)n, 0.0); d.assign((size_t)n, 0.0);
            for (int i = 0; i < n - 1; i++) {
                d[(size_t)i] = 1.0 / 3.0 * (c[(size_t)i + 1] - c[(size_t)i]) / (x[(size_t)i + 1] - x[(size_t)i]);
                b[(size_t)i] = (y[(size_t)i + 1] - y[(size_t)i]) / (x[(size_t)i + 1] - x[(size_t)i]) - 1.0 / 3.0 * (2.0 * c[(size_t)i] + c[(size_t)i + 1]) * (x[(size_t)i + 1] - x[(size_t)i]);
            }
            const double h = x[(size_t)n - 1] - x[(size_t)n - 2];
            d[(size_t)n - 1] = 0.0;
            b[(size_t)n - 1] = 3.0 * d[(size_t)n - 2] * h * h + 2.0 * c[(size_t)n - 2] * h + b[(size_t)n - 2];
            c0 = c[0];
            return true;
        }
        if (type == 2) {
            b.assign((size_t)n, 0.0); c.assign((size_t)n, 0.0); d.assign((size_t)n, 0.0);
            for (int i = 1; i < n - 1; i++) {
                const double h = x[(size_t)i + 1] - x[(size_t)i], hl = x[(size_t)i] - x[(size_t)i - 1];
                b[(size_t)i] = -h / (hl * (hl + h)) * y[(size_t)i - 1] + (h - hl) / (hl * h) * y[(size_t)i] + hl / (h * (hl + h)) * y[(size_t)i + 1];
            }
            { const double h = x[1] - x[0]; b[0] = 0.5 * (-b[1] - 0.5 * 0.0 * h + 3.0 * (y[1] - y[0]) / h); }
            {
                const double h = x[(size_t)n - 1] - x[(size_t)n - 2];
                b[(size_t)n - 1] = 0.5 * (-b[(size_t)n - 2] + 0.5 * 0.0 * h + 3.0 * (y[(size_t)n - 1] - y[(size_t)n - 2]) / h);
                c[(size_t)n - 1] = 0.5 * 0.0;
            }
            d[(size_t)n - 1] = 0.0;
            const double c_last = c[(size_t)n - 1];
            set_coeffs_from_b();
            c[(size_t)n - 1] = c_last; d[(size_t)n - 1] = 0.0;                     // set_coeffs_from_b leaves the last entries as set above
            return true;
        }
        return false;
    }
    double operator()(double xv) const                                              // spline.h:480-503
    {
        const size_t n = x.size();
        const long it = (long)(std::upper_bound(x.begin(), x.end(), xv) - x.begin());
        const size_t idx = (size_t)std::max(it - 1, 0L);
        const double h = xv - x[idx];
        if (xv < x[0]) return (c0 * h + b[0]) * h + y[0];
        if (xv > x[n - 1]) return (c[n - 1] * h + b[n - 1]) * h + y[n - 1];
        return ((d[idx] * h + c[idx]) * h + b[idx]) * h + y[idx];
    }
};

// Appourchaux et al. 2014 / 2016 width relation as the model functions evaluate it (models.cpp:4788-4794, 4974-4977)
inline double app_width(double f, const double g[6])
{
    const double lnGamma0 = g[2] * std::log(f / g[0]) + std::log(g[3]);
    const double e = 2. * std::log(f / g[1]) / std::log(g[4] / g[0]);
    const double lnLorentz = -std::log(g[5]) / (1. + std::pow(e, 2));
    return std::exp(lnGamma0 + lnLorentz);
}

}  // namespace

// ---------------------------------------------------------------------------------------------- prepare / solve / finish
// tamcmc_host_expand_rgb_v4 in three stages, so that the pair loop and the zeta normalisation -- 99 % of its time -- can run on the
// device (rgb_device.cu: tamcmc_gpu_rgb_expand) between the first and the last: everything below is the host code of both paths.
namespace tamcmc_rgb {

struct Prep {
    bool app = false;
    int Nmax = 0, lmax = 0, Nfl0 = 0, Nfl1 = 0, Nfl2 = 0, Nfl3 = 0, Nsplit = 0, Nwidth = 0, Nnoise = 0, Nf = 0;
    const double* params = nullptr;
    double trunc_c = 0, inclination = 0, Vl1 = 0, Vl2 = 0, Vl3 = 0, gp[6] = {0, 0, 0, 0, 0, 0};
    bool do_amp = false;
    double bias_type = 0, q_star = 0, Wfactor = 0, Hfactor = 0, rot_env = 0, rot_core = 0;
    double a2_env = 0, a3_env = 0, a4_env = 0, a5_env = 0, a6_env = 0, eta_switch = 0, asym = 0, fmin = 0, fmax = 0;
    vec fl0_all, Wl0_all, Hl0_all;
    Spline bias;
    Eigensols S;
    PairSetup setup;               // set when the solve was deferred
    vec fl1_all, ksi_raw;          // finish in stages: the mixed-mode frequencies (with bias) and the zeta sums at them
    bool have_ksi = false;
};

Prep* prep_new() { return new Prep(); }
void prep_free(Prep* p) { delete p; }

// Everything of model_RGB_asympt_aj_*Width_HarveyLike_v4 up to the mixed-mode solve (models.cpp:4700-4866 / 4350-4500).  With
// defer_solve the pair loop is NOT run: P->setup holds its arguments and P->S the p and g modes.
int prepare(Prep* Pp, int model_id, const double* params, const int* plength, double step, bool defer_solve)
{
    Prep& P = *Pp;
    if (!params || !plength) return TAMCMC_ERR_ARG;
    if (model_id != 25 && model_id != 27) return TAMCMC_ERR_MODEL;
    P.params = params;
    const bool app = P.app = (model_id == 25);
    const int Nmax = P.Nmax = plength[0], lmax = P.lmax = plength[1], Nfl0 = P.Nfl0 = plength[2], Nfl1 = P.Nfl1 = plength[3];
    P.Nfl2 = plength[4]; P.Nfl3 = plength[5];
    const int Nsplit = P.Nsplit = plength[6], Nwidth = P.Nwidth = plength[7], Nnoise = P.Nnoise = plength[8], Ninc = plength[9], Ncfg = plength[10];
    const int Nf = P.Nf = Nfl0 + Nfl1 + P.Nfl2 + P.Nfl3;
    const int o_cfg = Nmax + lmax + Nf + Nsplit + Nwidth + Nnoise + Ninc;
    if (Nmax < 2 || Nfl0 != Nmax || lmax < 1 || lmax > 3 || Nsplit < 10 || Nnoise < 1 || Ncfg < 6 || Nwidth < (app ? 6 : 1) || Nfl1 < 8) return TAMCMC_ERR_ARG;
    P.trunc_c = params[o_cfg];
    P.do_amp = params[o_cfg + 1] != 0.0;
    const double model_type = params[o_cfg + 3];
    P.bias_type = params[o_cfg + 4];
    const int Nferr = (int)params[o_cfg + 5];
    if (Nferr < 0 || Nfl1 < 8 + 2 * Nferr) return TAMCMC_ERR_ARG;
    const ld pi = M_PI;
    const int o_w = Nmax + lmax + Nf + Nsplit;
    if (app) for (int k = 0; k < 6; k++) P.gp[k] = std::abs(params[o_w + k]);
    P.inclination = std::abs(params[Nmax + lmax + Nf + Nsplit + Nwidth + Nnoise]);
    P.Vl1 = std::abs(params[Nmax]);
    P.Vl2 = (lmax >= 2) ? std::abs(params[Nmax + 1]) : 0.0; P.Vl3 = (lmax >= 3) ? std::abs(params[Nmax + 2]) : 0.0;

    // ---- l = 0: frequencies, widths, heights (models.cpp:4784-4803 / 4433-4442) ----
    P.fl0_all.assign(params + Nmax + lmax, params + Nmax + lmax + Nfl0);
    const vec& fl0_all = P.fl0_all;
    P.Wl0_all.resize((size_t)Nmax); P.Hl0_all.resize((size_t)Nmax);
    for (int n = 0; n < Nmax; n++) P.Wl0_all[(size_t)n] = app ? app_width(fl0_all[(size_t)n], P.gp) : std::abs(params[o_w]);
    for (int n = 0; n < Nmax; n++)
        P.Hl0_all[(size_t)n] = P.do_amp ? std::abs(params[n] * ((1.0 / P.Wl0_all[(size_t)n]) / (double)pi)) : std::abs(params[n]);

    // ---- mixed modes (models.cpp:4811-4866) ----
    const int o_l1 = Nmax + lmax + Nfl0;
    const double delta0l = params[o_l1], DPl = std::abs(params[o_l1 + 1]), alpha_g = std::abs(params[o_l1 + 2]);
    const double q_star = P.q_star = std::abs(params[o_l1 + 3]);
    P.Wfactor = std::abs(params[o_l1 + 6]); P.Hfactor = std::abs(params[o_l1 + 7]);
    const int o_s = Nmax + lmax + Nf;
    P.rot_env = std::abs(params[o_s]); P.rot_core = std::abs(params[o_s + 1]);
    P.a2_env = params[o_s + 2]; P.a3_env = params[o_s + 4]; P.a4_env = params[o_s + 5]; P.a5_env = params[o_s + 6]; P.a6_env = params[o_s + 7];
    P.eta_switch = params[o_s + 8]; P.asym = params[o_s + 9];
    vec fref_all((size_t)Nferr), ferr_all((size_t)Nferr);
    for (int i = 0; i < Nferr; i++) { fref_all[(size_t)i] = params[o_l1 + 8 + i]; ferr_all[(size_t)i] = params[o_l1 + 8 + Nferr + i]; }
    if (P.bias_type != 0) {
        if (P.bias_type != 1 && P.bias_type != 2) return TAMCMC_ERR_MODEL;      // (the reference would evaluate an unset spline)
        if (!P.bias.set_points(fref_all, ferr_all, (int)P.bias_type)) return TAMCMC_ERR_ARG;      // tk::spline asserts >= 3 strictly increasing points
    }
    const double fmin = P.fmin = vmin(fl0_all), fmax = P.fmax = vmax(fl0_all);
    PairSetup* defer = defer_solve ? &P.setup : nullptr;
    P.setup = PairSetup();
    if (app || model_type == 0) {
        double rfit[2];
        linfit(linspaced(Nmax, 0.0, (double)(Nmax - 1)), fl0_all, rfit);
        const double Dnu_p = rfit[0];
        const int n0 = (int)std::floor(rfit[1] / Dnu_p);
        const double epsilon_p = rfit[1] / Dnu_p - n0;
        if (fmin - Dnu_p < 0) return TAMCMC_ERR_NONFINITE;                       // "THE ARMM WILL NOT CONVERGE": the reference exits (models.cpp:4852-4858)
        if (model_type == 0) {
            const int rc = solve_O2p(Dnu_p, epsilon_p, 1, delta0l, 0, 0., DPl, alpha_g, q_star, fmin - Dnu_p, fmax + Dnu_p, step, P.S, defer);
            if (rc) return rc;
        }
    }
    if (model_type != 0) solve_from_l0(fl0_all, 1, delta0l, DPl, alpha_g, q_star, step, fmin, fmax, P.S, defer);
    if (!P.S.ok) return TAMCMC_ERR_NONFINITE;
    return TAMCMC_OK;
}

// First part of the rest: the mixed-mode frequencies (deferred: filter, sort, unique of the raw solutions; then the bias, models.cpp:4868-4874).
int finish_modes(Prep* Pp, bool deferred, const double* cand, int ncand)
{
    Prep& P = *Pp;
    Eigensols& S = P.S;
    P.have_ksi = false;
    if (deferred) {
        if (!P.setup.set) return TAMCMC_ERR_NONFINITE;
        vec all;
        if (cand && ncand > 0) all.assign(cand, cand + ncand);
        sort_unique(all, P.setup.resol, P.setup.keep_min, P.setup.keep_max, S.nu_m);
    }
    if (!S.ok || S.nu_m.empty()) return TAMCMC_ERR_NONFINITE;                    // no mixed mode: the reference indexes empty vectors from here on
    P.fl1_all = S.nu_m;
    if (P.bias_type != 0) for (double& f : P.fl1_all) f = f + P.bias(f);
    P.ksi_raw.assign(P.fl1_all.size(), 0.0);
    return TAMCMC_OK;
}
// The zeta sums at the mixed modes, one block of 8 frequencies at a time (so that a caller can spread the blocks of all its chains over
// its threads): ksi_blocks() blocks, any order, then ksi_done().
int ksi_blocks(const Prep* P) { return (int)((P->fl1_all.size() + KSI_W - 1) / KSI_W); }
void ksi_block_compute(Prep* Pp, int b)
{
    Prep& P = *Pp;
    const KsiConst K(P.S.nu_g, P.S.dnup, P.q_star);
    KsiScratch W;
    ksi_sum_block(b, P.fl1_all, P.S.nu_p, P.S.dnup, P.S.dPg, K, W, P.ksi_raw);
}
void ksi_done(Prep* P) { P->have_ksi = true; }
// The exact zeta sums (host arithmetic, host cosines) at points `idx` of the 4-year-resolution grid, and their maximum: the device finds
// WHERE the maximum is with its fast arithmetic, the value that normalises the zeta function is computed here like the reference does.
double ksi_norm_at(const Prep* Pp, const int* idx, int n)
{
    const Prep& P = *Pp;
    double f0, f1; int Ndata;
    if (n < 1 || !ksi_highres_grid(P.S.nu_p, P.S.nu_g, f0, f1, Ndata)) return -1.0;
    vec nu((size_t)n), out;
    const double st = (Ndata > 1) ? (f1 - f0) / (double)(Ndata - 1) : 0.0;
    for (int k = 0; k < n; k++) {
        if (idx[k] < 0 || idx[k] >= Ndata) return -1.0;
        nu[(size_t)k] = (Ndata == 1 || idx[k] == Ndata - 1) ? f1 : f0 + (double)idx[k] * st;          // linspaced(Ndata, f0, f1)[idx]
    }
    const KsiConst K(P.S.nu_g, P.S.dnup, P.q_star);
    KsiScratch W;
    out.assign((size_t)n, 0.0);
    for (long b = 0; b < (long)((n + KSI_W - 1) / KSI_W); b++) ksi_sum_block(b, nu, P.S.nu_p, P.S.dnup, P.S.dPg, K, W, out);
    return vmax(out);
}

// The rest of the model function (models.cpp:4868-5006): bias, zeta function, heights / widths / splittings of the mixed modes, the row.
// deferred: cand / ncand are the raw solutions of a deferred pair loop (any order; filtered, sorted and made unique here).
// norm: max of the zeta sums over the 4-year-resolution grid (bump_DP.cpp:155-163) when it was computed elsewhere, < 0 to compute it here.
int finish(Prep* Pp, bool deferred, const double* cand, int ncand, double norm, int capacity, double* row_out, int* nmodes_out)
{
    Prep& P = *Pp;
    const double* params = P.params;
    const bool app = P.app;
    const int Nmax = P.Nmax, lmax = P.lmax, Nfl0 = P.Nfl0, Nfl1 = P.Nfl1, Nfl2 = P.Nfl2, Nfl3 = P.Nfl3, Nsplit = P.Nsplit, Nwidth = P.Nwidth, Nnoise = P.Nnoise, Nf = P.Nf;
    const ld pi = M_PI;
    const vec& fl0_all = P.fl0_all; const vec& Wl0_all = P.Wl0_all; const vec& Hl0_all = P.Hl0_all;
    Eigensols& S = P.S;
    if (!P.have_ksi) {
        const int rc = finish_modes(Pp, deferred, cand, ncand);
        if (rc) return rc;
    }
    const vec& fl1_all = P.fl1_all;
    vec ksi_pg = P.ksi_raw;
    const bool have = P.have_ksi;
    P.have_ksi = false;
    if (!ksi_fct2_precise(fl1_all, S.nu_p, S.nu_g, S.dnup, S.dPg, P.q_star, ksi_pg, norm, have)) return TAMCMC_ERR_NONFINITE;
    const size_t N1 = fl1_all.size();
    const double Hfactor = P.Hfactor, Wfactor = P.Wfactor, rot_env = P.rot_env, rot_core = P.rot_core, fmin = P.fmin, fmax = P.fmax;
    const double Vl1 = P.Vl1, Vl2 = P.Vl2, Vl3 = P.Vl3;
    // h_l_rgb (bump_DP.cpp:235-253)
    vec h1_h0((size_t)N1);
    for (size_t i = 0; i < N1; i++) {
        double v = std::sqrt(1.0 - (double)(ld)Hfactor * ksi_pg[i]);
        if (v > 0 - 1e-5 && v < 0 + 1e-5) v = 1e-10;
        h1_h0[i] = v;
    }
    // heights of the l=1 modes: l=0 heights interpolated with zero anchors outside the comb (models.cpp:4879-4905)
    vec f_interp((size_t)Nmax + 4), h_interp((size_t)Nmax + 4);
    f_interp[0] = fmin * 0.6; f_interp[1] = fmin * 0.8; f_interp[(size_t)Nmax + 2] = fmax * 1.2; f_interp[(size_t)Nmax + 3] = fmax * 1.4;
    h_interp[0] = 0; h_interp[1] = Hl0_all[0] / 4; h_interp[(size_t)Nmax + 2] = Hl0_all[(size_t)Nmax - 1] / 4; h_interp[(size_t)Nmax + 3] = 0;
    for (int j = 0; j < Nmax; j++) { f_interp[(size_t)j + 2] = fl0_all[(size_t)j]; h_interp[(size_t)j + 2] = Hl0_all[(size_t)j]; }
    vec Hl1_all(N1), Wl1_all(N1), a1_l1(N1);
    for (size_t i = 0; i < N1; i++) {
        const double tmp = tamcmc_host::lin_interpol(f_interp.data(), h_interp.data(), Nmax + 4, fl1_all[i]);
        const double Hl1p = (tmp < 0) ? 0.0 : std::abs(tmp);
        Hl1_all[i] = h1_h0[i] * (Hl1p * Vl1);
        // gamma_l_fct2 (bump_DP.cpp:203-223): long double arithmetic around the interpolated l=0 width
        const ld width0_at_l = tamcmc_host::lin_interpol(fl0_all.data(), Wl0_all.data(), Nmax, fl1_all[i]);
        Wl1_all[i] = (double)(width0_at_l * (1. - (ld)Wfactor * ksi_pg[i]) / std::sqrt(h1_h0[i]));
        // dnu_rot_2zones (bump_DP.cpp:531-537), then abs (models.cpp:4909)
        a1_l1[i] = std::abs(ksi_pg[i] * (double)((ld)rot_core / 2 - (ld)rot_env) + (double)(ld)rot_env);
    }
    const double eta0 = (P.eta_switch == 1) ? tamcmc_host::eta0_fct(fl0_all.data(), Nmax) : 0.0;

    // ---- the row: header, noise, then one record per optimum_lorentzian_calc_aj call (models.cpp:4931-5006) ----
    const int nmodes = Nfl0 + (int)N1 + Nfl2 + Nfl3;
    if (nmodes_out) *nmodes_out = nmodes;
    if (nmodes > capacity) return TAMCMC_ERR_ARG;
    const int row_len = TAMCMC_MT_HEADER + Nnoise + TAMCMC_MT_STRIDE * capacity;
    std::memset(row_out, 0, sizeof(double) * (size_t)row_len);
    row_out[0] = nmodes; row_out[1] = P.inclination; row_out[2] = P.trunc_c; row_out[3] = P.asym;
    for (int k = 0; k < Nnoise; k++) row_out[TAMCMC_MT_HEADER + k] = params[Nmax + lmax + Nf + Nsplit + Nwidth + k];
    double* rec = row_out + TAMCMC_MT_HEADER + Nnoise;
    int j = 0;
    for (int n = 0; n < Nfl0; n++, j++) { double* r = rec + (size_t)TAMCMC_MT_STRIDE * j; r[0] = 0; r[1] = fl0_all[(size_t)n]; r[2] = Hl0_all[(size_t)n]; r[3] = Wl0_all[(size_t)n]; }
    for (size_t n = 0; n < N1; n++, j++) {
        double* r = rec + (size_t)TAMCMC_MT_STRIDE * j;
        r[0] = 1; r[1] = fl1_all[n]; r[2] = std::abs(Hl1_all[n]); r[3] = Wl1_all[n]; r[4] = a1_l1[n]; r[10] = eta0;
    }
    for (int l = 2; l <= 3; l++) {
        const int Nfl = (l == 2) ? Nfl2 : Nfl3;
        const int off = Nmax + lmax + Nfl0 + Nfl1 + ((l == 3) ? Nfl2 : 0);
        const double Vl = (l == 2) ? Vl2 : Vl3;
        for (int n = 0; n < Nfl; n++, j++) {
            double* r = rec + (size_t)TAMCMC_MT_STRIDE * j;
            const double fl = std::abs(params[off + n]);
            if (!app && n >= Nmax) return TAMCMC_ERR_ARG;                        // the CteWidth model reads Wl0_all[n] (models.cpp:4597, 4617)
            const double W = app ? app_width(fl, P.gp) : Wl0_all[(size_t)n];
            const double Hi = tamcmc_host::lin_interpol(fl0_all.data(), Hl0_all.data(), Nmax, fl);
            const double H = P.do_amp ? (double)fabsl(Hi / (pi * W) * Vl) : std::abs(Hi * Vl);
            r[0] = l; r[1] = fl; r[2] = H; r[3] = W; r[4] = rot_env; r[5] = P.a2_env; r[6] = P.a3_env; r[7] = P.a4_env;
            if (l == 3) { r[8] = P.a5_env; r[9] = P.a6_env; }
            r[10] = eta0;
        }
    }
    return TAMCMC_OK;
}

// What the device solver needs for one chain (rgb_device.cu), appended to T: one Band per p mode (coarse grid, band, constants of p - g
// and of the local grids), one Pair per (p mode, g mode) whose g mode lies in the p mode's search range (solver_mm.cpp:343), and the sums
// of the zeta normalisation.  false: solve on the host.
bool export_task(const Prep* Pp, int chain, DeviceTask& T)
{
    const Prep& P = *Pp;
    const PairSetup& U = P.setup;
    const Eigensols& S = P.S;
    if (!U.set || !g_armm_fast || S.nu_p.empty() || S.nu_g.empty()) return false;
    const PminusG probe(S.nu_p[0], S.nu_g[0], U.dnu_local[0], U.DPl, U.q);
    if (!probe.usable()) return false;
    double kf0, kf1; int Ndata;
    if (!ksi_highres_grid(S.nu_p, S.nu_g, kf0, kf1, Ndata)) return false;
    for (double d : S.dPg) if (d != S.dPg[0]) return false;                // (the device's zeta kernel relies on the common period spacing)
    const ld D = U.resol * (ld)U.fact;                                   // the local grids' step, as solver_mm forms it (solver_mm.cpp:406)
    const double Dh = (double)D, Dl = (double)(D - (ld)Dh);
    const size_t bands0 = T.bands.size();
    for (size_t np = 0; np < S.nu_p.size(); np++) {
        const BandGrid G = band_of(S.nu_p[np], U.dnu_local[np], U.lo[np], U.hi[np], U.resol);
        Band B;
        B.nu_p = (double)(ld)S.nu_p[np]; B.Dnu = (double)(ld)U.dnu_local[np];
        B.lo = G.lo; B.hi = G.hi; B.gstep = G.gstep; B.DPl = (double)U.DPl; B.q = (double)U.q;
        B.resol2 = (double)(2 * U.resol); B.Dh = Dh; B.Dl = Dl;
        B.n = (int)G.n; B.i_lo = (int)G.i_lo; B.nband = G.valid ? (int)(G.i_hi - G.i_lo + 1) : 0; B.slot_off = T.nslots; B.chain = chain; B.nseg_est = 1; B.rep_inv_g = 0.0;
        if (!(U.dnu_local[np] > 0) || G.n > 2000000000L) return false;
        if (B.nband > 0 && B.nband < 8) return false;                    // a band of a few points: the host's full scan
        T.nslots += B.nband;
        int first = 1;
        for (size_t ng = 0; ng < S.nu_g.size(); ng++) {
            const ld nu_g = S.nu_g[ng];
            if (!(nu_g >= U.lo[np] && nu_g <= U.hi[np]) || B.nband == 0) continue;
            Pair Q; Q.inv_g = 1.0 / (double)nu_g; Q.nu_g = (double)nu_g; Q.band = (int)(bands0 + np); Q.pad_ = 0;
            if (first) {                                                 // segments of a pair of this band = poles of the tangent inside it + 1
                const PminusG F(S.nu_p[np], nu_g, U.dnu_local[np], U.DPl, U.q);
                const double npoles = std::floor(F.u_of(G.grid(G.i_lo)) - 0.5) - std::ceil(F.u_of(G.grid(G.i_hi)) - 0.5) + 1.0;
                B.nseg_est = (npoles >= 0.0 && npoles < 1.0e6) ? (int)npoles + 2 : 1;
                B.rep_inv_g = Q.inv_g;
                first = 0;
            }
            T.pairs.push_back(Q);
        }
        T.bands.push_back(B);
    }
    KsiHdr K;
    const ld pi = M_PI;
    K.fmin = kf0; K.fmax = kf1; K.c_up = (double)(pi * 1e6); K.pi_d = (double)pi;
    K.Lp = (int)S.nu_p.size(); K.Lg = (int)S.nu_g.size(); K.Ndata = Ndata; K.off_p = (int)(T.kp.size() / 3); K.off_g = (int)(T.kg.size() / 2); K.chain = chain; K.val_off = T.nvals; K.pad_ = 0;
    T.nvals += Ndata;
    for (size_t p = 0; p < S.nu_p.size(); p++) { T.kp.push_back(S.nu_p[p]); T.kp.push_back(S.dnup[p]); T.kp.push_back((double)(U.q * (ld)S.dnup[p])); }
    for (size_t g = 0; g < S.nu_g.size(); g++) { T.kg.push_back((double)(1. / (ld)S.nu_g[g])); T.kg.push_back(S.dPg[g]); }
    T.ksi.push_back(K);
    return true;
}

// TEST HOOK: the reference's long double local grid (local_grid above) for the emulation's self-check
long local_grid_reference(const Prep* P, double nu_idx, double* lo, double* hi)
{
    return local_grid(nu_idx, P->setup.resol, (ld)P->setup.fact, *lo, *hi);
}

}  // namespace tamcmc_rgb

extern "C" {

// Replaces: solve_mm_asymptotic_O2from_l0 (external/ARMM/solver_mm.cpp:624-746) with sigma_p = 0, returns_pg_freqs = true.
// Arrays of capacity `cap`; counts in n_m, n_p, n_g.  Returns TAMCMC_ERR_ARG when a capacity is too small, TAMCMC_ERR_NONFINITE
// where the reference gives up ("impossible star", no p mode in range).
int tamcmc_host_armm_solve_from_l0(const double* nu_l0, int n_l0, int el, double delta0l, double DPl, double alpha, double q, double resol,
                                   double freq_min, double freq_max, int cap, double* nu_m, int* n_m, double* nu_p, double* dnup, int* n_p,
                                   double* nu_g, int* n_g)
{
    if (!nu_l0 || n_l0 < 2 || !nu_m || !n_m || cap < 1) return TAMCMC_ERR_ARG;
    Eigensols S;
    solve_from_l0(vec(nu_l0, nu_l0 + n_l0), el, delta0l, DPl, alpha, q, resol, freq_min, freq_max, S);
    if (!S.ok) return TAMCMC_ERR_NONFINITE;
    if ((int)S.nu_m.size() > cap || (int)S.nu_p.size() > cap || (int)S.nu_g.size() > cap) return TAMCMC_ERR_ARG;
    *n_m = (int)S.nu_m.size(); std::copy(S.nu_m.begin(), S.nu_m.end(), nu_m);
    if (n_p) *n_p = (int)S.nu_p.size();
    if (nu_p) std::copy(S.nu_p.begin(), S.nu_p.end(), nu_p);
    if (dnup) std::copy(S.dnup.begin(), S.dnup.end(), dnup);
    if (n_g) *n_g = (int)S.nu_g.size();
    if (nu_g) std::copy(S.nu_g.begin(), S.nu_g.end(), nu_g);
    return TAMCMC_OK;
}

// Replaces: solve_mm_asymptotic_O2p (external/ARMM/solver_mm.cpp:470-604) with sigma_p = 0, returns_pg_freqs = true.
int tamcmc_host_armm_solve_O2p(double Dnu_p, double epsilon, int el, double delta0l, double alpha_p, double nmax, double DPl, double alpha,
                               double q, double fmin, double fmax, double resol, int cap, double* nu_m, int* n_m, double* nu_p, double* dnup,
                               int* n_p, double* nu_g, int* n_g)
{
    if (!nu_m || !n_m || cap < 1) return TAMCMC_ERR_ARG;
    Eigensols S;
    const int rc = solve_O2p(Dnu_p, epsilon, el, delta0l, alpha_p, nmax, DPl, alpha, q, fmin, fmax, resol, S);
    if (rc) return rc;
    if (!S.ok) return TAMCMC_ERR_NONFINITE;
    if ((int)S.nu_m.size() > cap || (int)S.nu_p.size() > cap || (int)S.nu_g.size() > cap) return TAMCMC_ERR_ARG;
    *n_m = (int)S.nu_m.size(); std::copy(S.nu_m.begin(), S.nu_m.end(), nu_m);
    if (n_p) *n_p = (int)S.nu_p.size();
    if (nu_p) std::copy(S.nu_p.begin(), S.nu_p.end(), nu_p);
    if (dnup) std::copy(S.dnup.begin(), S.dnup.end(), dnup);
    if (n_g) *n_g = (int)S.nu_g.size();
    if (nu_g) std::copy(S.nu_g.begin(), S.nu_g.end(), nu_g);
    return TAMCMC_OK;
}

// Replaces: tk::spline set_boundary(second_deriv, 0, second_deriv, 0) + set_points(x, y, cspline | cspline_hermite) + operator()
// (external/spline/src/spline.h), as the bias of the l=1 mixed modes uses it (models.cpp:4834-4843, 4870-4874).  type: 1 cubic, 2 Hermite.
int tamcmc_host_spline_eval(const double* x, const double* y, int n, int type, const double* xq, int nq, double* out)
{
    if (!x || !y || !xq || !out || n < 3 || nq < 0) return TAMCMC_ERR_ARG;
    Spline s;
    if (!s.set_points(vec(x, x + n), vec(y, y + n), type)) return TAMCMC_ERR_ARG;
    for (int i = 0; i < nq; i++) out[i] = s(xq[i]);
    return TAMCMC_OK;
}

// Replaces: the host half of model_RGB_asympt_aj_AppWidth_HarveyLike_v4 (model_id 25, models.cpp:4684-4927) and of
// model_RGB_asympt_aj_CteWidth_HarveyLike_v4 (model_id 27, models.cpp:4334-4556): params / plength in those models' layout
//   params = [H(Nmax) | V_l(lmax) | fl0(Nfl0) | l=1 block(Nfl1) = delta0l, DPl, alpha_g, q, -, -, Wfactor, Hfactor, fref[Nferr], ferr[Nferr]
//             | fl2 | fl3 | split(Nsplit) = rot_env, rot_core, a2_env, a2_core, a3_env, a4_env, a5_env, a6_env, eta_switch, asym
//             | width(Nwidth) | noise(Nnoise) | inc | cfg = trunc_c, do_amp, sigma_limit, model_type, bias_type, Nferr]
// -> ONE mode-table row of `capacity` modes (row_out: TAMCMC_MT_HEADER + Nnoise + TAMCMC_MT_STRIDE * capacity doubles), modes in the
// reference's call order (all l=0, the mixed l=1 modes, l=2, l=3).  step = x[2] - x[1] of the spectrum (models.cpp:4714): the
// resolution of the mixed-mode solver's grid.  *nmodes_out receives the number of modes (the l=1 count varies from chain to chain);
// TAMCMC_ERR_ARG with *nmodes_out set when it exceeds `capacity`.
int tamcmc_host_expand_rgb_v4(int model_id, const double* params, const int* plength, double step, int capacity, double* row_out,
                              int* nmodes_out)
{
    if (!params || !plength || !row_out || capacity < 1) return TAMCMC_ERR_ARG;
    tamcmc_rgb::Prep P;
    int rc = tamcmc_rgb::prepare(&P, model_id, params, plength, step, /*defer_solve=*/false);
    if (rc) return rc;
    return tamcmc_rgb::finish(&P, false, nullptr, 0, -1.0, capacity, row_out, nmodes_out);
}

}  // extern "C"
