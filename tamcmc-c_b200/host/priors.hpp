// priors.hpp -- host-side generic priors for the C++ driver (host/mcmc_driver.hpp): the primitive log-probabilities of
// tamcmc/sources/stats_dictionary.cpp and their dispatch per parameter (apply_generic_priors, priors_calc.cpp:725-860,
// switch values = Config/default/primepriors_ctrl.list), plus the two prior functions of the Gaussian-envelope models
// (priors_Harvey_Gaussian / priors_Kallinger2014_Gaussian, priors_calc.cpp:631-700).  Priors stay on the host
// (SURVEY.md 8b): a -inf prior masks the chain out of the batched GPU evaluation (model_def.cpp:469, 476-480).
//
// The reference computes these in long double; so does this header (the sums feed accept/reject decisions at O(1)).
// Tabulated priors (cases 11, 12: GSL 2-D interpolators) and the multivariate Gaussian (case 3: the reference exits)
// are not offered: GenericPriors::valid() reports them.
#pragma once
#include <cmath>
#include <cstddef>
#include <limits>
#include <string>
#include <vector>

namespace tamcmc {
namespace priors {

constexpr long double PIl = 3.141592653589793238462643383279502884L;
constexpr long double NEG_INF = -std::numeric_limits<long double>::infinity();

// primepriors_ctrl.list
enum Kind { NONE = 0, UNIFORM = 1, GAUSSIAN = 2, MULTIVAR_GAUSSIAN = 3, JEFFREYS = 4, UG = 5, GU = 6, GUG = 7, UNIFORM_ABS = 8,
            UNIFORM_COS = 9, JEFFREYS_ABS = 10, TABULATED = 11, TABULATED_2D = 12, AUTO = 13 };

inline int kind_of(const std::string& name)
{
    static const char* names[] = {"None", "Uniform", "Gaussian", "multivar_Gaussian", "Jeffreys", "UG", "GU", "GUG", "Uniform_abs",
                                  "Uniform_cos", "Jeffreys_abs", "Tabulated", "Tabulated_2d", "Auto"};
    if (name == "Fix") return NONE;
    for (int k = 0; k < 14; k++) if (name == names[k]) return k;
    return -1;
}

// stats_dictionary.cpp:38-52
inline long double logP_uniform(long double b_min, long double b_max, long double x)
{
    return (x <= b_max && x >= b_min) ? -std::log(std::fabs(b_max - b_min)) : NEG_INF;
}
// stats_dictionary.cpp:56-70
inline long double logP_uniform_abs(long double b_min, long double b_max, long double x)
{
    return (std::fabs(x) <= b_max && std::fabs(x) >= b_min) ? -std::log(std::fabs(b_max - b_min)) : NEG_INF;
}
// stats_dictionary.cpp:74-94 (x in degrees; strict bounds, no Jacobian -- "Buggy. Do not use" in the control list, kept as is)
inline long double logP_uniform_cos(long double b_min, long double b_max, long double x)
{
    const long double c = std::cos(PIl * x / 180.L);
    return (c < b_max && c > b_min) ? -std::log(std::fabs(b_max - b_min)) : NEG_INF;
}
// stats_dictionary.cpp:98-106
inline long double logP_gaussian(long double mean, long double sigma, long double x)
{
    const long double z = (x - mean) / sigma;
    return -std::log(std::sqrt(2 * PIl) * sigma) - 0.5L * (z * z);
}
// stats_dictionary.cpp:127-145: truncated Jeffreys 1/(h + hmin) on (0, hmax)
inline long double logP_jeffrey(long double hmin, long double hmax, long double h)
{
    if (!(h < hmax && h > 0)) return NEG_INF;
    const long double prior = 1.L / (h + hmin), norm = std::log((hmax + hmin) / hmin);
    return std::log(prior / norm);
}
// stats_dictionary.cpp:149-167
inline long double logP_jeffrey_abs(long double hmin, long double hmax, long double h)
{
    if (!(std::fabs(h) < hmax)) return NEG_INF;
    const long double prior = 1.L / (std::fabs(h) + hmin), norm = std::log((hmax + hmin) / hmin);
    return std::log(prior / norm);
}
// stats_dictionary.cpp:173-196: flat on [b_min, b_max], Gaussian wing above
inline long double logP_uniform_gaussian(long double b_min, long double b_max, long double sigma, long double x)
{
    long double logP = std::numeric_limits<long double>::quiet_NaN();     // the reference leaves logP unset for NaN inputs
    if (x < b_min) logP = NEG_INF;
    if (x <= b_max && x >= b_min) logP = 0;
    if (x > b_max) { const long double z = (x - b_max) / sigma; logP = -0.5L * (z * z); }
    return logP - std::log(std::fabs(b_max - b_min) + 0.5L * std::sqrt(2 * PIl) * sigma);
}
// stats_dictionary.cpp:200-222: Gaussian wing below, flat on [b_min, b_max]
inline long double logP_gaussian_uniform(long double b_min, long double b_max, long double sigma, long double x)
{
    long double logP = std::numeric_limits<long double>::quiet_NaN();
    if (x > b_max) logP = NEG_INF;
    if (x <= b_max && x >= b_min) logP = 0;
    if (x < b_min) { const long double z = (x - b_min) / sigma; logP = -0.5L * (z * z); }
    return logP - std::log(std::fabs(b_max - b_min) + 0.5L * std::sqrt(2 * PIl) * sigma);
}
// stats_dictionary.cpp:226-248
inline long double logP_gaussian_uniform_gaussian(long double b_min, long double b_max, long double sigma1, long double sigma2, long double x)
{
    long double logP = std::numeric_limits<long double>::quiet_NaN();
    if (x < b_min) { const long double z = (x - b_min) / sigma1; logP = -0.5L * (z * z); }
    if (x <= b_max && x >= b_min) logP = 0;
    if (x > b_max) { const long double z = (x - b_max) / sigma2; logP = -0.5L * (z * z); }
    return logP - std::log(std::fabs(b_max - b_min) + 0.5L * std::sqrt(2 * PIl) * (sigma1 + sigma2));
}

// One prior per parameter: `priors_params` is the reference's MatrixXd(4, Nparams) (Input_Data.priors, data.h:58; -9999 in
// unused slots), stored here row-major [4][Nparams]; `kinds` is priors_names_switch.
struct GenericPriors {
    std::vector<int> kinds;
    std::vector<double> p[4];

    GenericPriors() {}
    GenericPriors(const std::vector<int>& kinds_, const std::vector<double>& p0, const std::vector<double>& p1,
                  const std::vector<double>& p2, const std::vector<double>& p3) : kinds(kinds_) { p[0] = p0; p[1] = p1; p[2] = p2; p[3] = p3; }

    // false when a prior kind needs what this header does not carry (tables, the multivariate Gaussian) or is unknown
    bool valid() const
    {
        for (int k : kinds) if (k < 0 || k == MULTIVAR_GAUSSIAN || k == TABULATED || k == TABULATED_2D || k > AUTO) return false;
        return true;
    }

    // apply_generic_priors (priors_calc.cpp:725-860): sum over the parameters, in parameter order
    long double apply(const double* params) const
    {
        long double pena = 0;
        for (size_t i = 0; i < kinds.size(); i++) {
            const long double a = p[0][i], b = p[1][i], c = p[2][i], d = p[3][i], x = params[i];
            switch (kinds[i]) {
            case UNIFORM:      pena = pena + logP_uniform(a, b, x); break;
            case GAUSSIAN:     pena = pena + logP_gaussian(a, b, x); break;
            case JEFFREYS:     pena = pena + logP_jeffrey(a, b, x); break;
            case UG:           pena = pena + logP_uniform_gaussian(a, b, c, x); break;
            case GU:           pena = pena + logP_gaussian_uniform(a, b, c, x); break;
            case GUG:          pena = pena + logP_gaussian_uniform_gaussian(a, b, c, d, x); break;
            case UNIFORM_ABS:  pena = pena + logP_uniform_abs(a, b, x); break;
            case UNIFORM_COS:  pena = pena + logP_uniform_cos(a, b, x); break;
            case JEFFREYS_ABS: pena = pena + logP_jeffrey_abs(a, b, x); break;
            default: break;                      // NONE / Fix / Auto: no prior applied
            }
        }
        return pena;
    }
};

// priors_Harvey_Gaussian (priors_calc.cpp:631-647): the envelope may not be narrower than half the large separation expected
// from numax (Stello+2009: Dnu = 0.263 numax^0.77); params[8] = numax, params[9] = sigma
inline long double priors_Harvey_Gaussian(const double* params, const GenericPriors& g)
{
    const long double Dnu_expected = 0.263L * std::pow((long double)params[8], 0.77L);
    if (params[9] < Dnu_expected / 2) return NEG_INF;
    return g.apply(params);
}

// priors_Kallinger2014_Gaussian (priors_calc.cpp:649-700); params[15..18] = numax, sigma, mu_numax, omega_numax
inline long double priors_Kallinger2014_Gaussian(const double* params, const GenericPriors& g)
{
    const long double numax = params[15], sig_numax = params[16], mu_numax = params[17], omega_numax = params[18];
    const long double Dnu_expected = 0.263L * std::pow(numax, 0.77L);
    if (params[5] < 0 || params[6] < 0) return NEG_INF;
    if (sig_numax < Dnu_expected / 2) return NEG_INF;
    if (numax + mu_numax < 0) return NEG_INF;
    long double f = logP_gaussian(0, std::fabs(omega_numax), mu_numax);
    f = f + g.apply(params);
    return f;
}

// 1-D tabulated prior (switch value 11), logP_tabulated (stats_dictionary.cpp:252-291) on a table read by
// formats.read_tabulated_prior / the reference's `.priors` files: linear interpolation of the PDF (interpol.cpp:13-43), a
// negative interpolated value counts as 0, outside the table the reference returns numeric_limits<double>::lowest().  The flag
// keeps the reference's meaning: normalise == false divides by the trapezoid area of the table, normalise == true divides by
// C = 0 exactly as the reference does (log(P) - log(0) = +inf: the caller is expected to pass false).
inline long double logP_tabulated(const double* tab_x, const double* tab_y, int n, long double x, bool normalise)
{
    double mn = tab_x[0], mx = tab_x[0];
    for (int i = 1; i < n; i++) { mn = tab_x[i] < mn ? tab_x[i] : mn; mx = tab_x[i] > mx ? tab_x[i] : mx; }
    if (x < mn || x > mx) return std::numeric_limits<double>::lowest();
    // lin_interpol(tab_x, tab_y, x): the reference's function takes the abscissa as a double
    const double xi = (double)x;
    int i = 0;
    double a = 0, b = 0;
    if (xi >= tab_x[0] && xi <= tab_x[n - 1]) {
        while ((xi < tab_x[i] || xi > tab_x[i + 1]) && i < n - 2) i = i + 1;
        a = (tab_y[i + 1] - tab_y[i]) / (tab_x[i + 1] - tab_x[i]);
        b = tab_y[i] - a * tab_x[i];
    }
    if (xi < tab_x[0]) { a = (tab_y[1] - tab_y[0]) / (tab_x[1] - tab_x[0]); b = tab_y[0] - a * tab_x[0]; }
    if (xi > tab_x[n - 1]) { a = (tab_y[n - 1] - tab_y[n - 2]) / (tab_x[n - 1] - tab_x[n - 2]); b = tab_y[n - 2] - a * tab_x[n - 2]; }
    long double P = a * xi + b;
    if (P < 0) P = 0;
    long double C = 0;
    if (!normalise)
        for (int k = 0; k + 1 < n; k++) { const double dy = (tab_y[k] + tab_y[k + 1]) / 2; const double dx = tab_x[k + 1] - tab_x[k]; C = C + dx * dy; }
    return std::log(P) - std::log(C);
}

}  // namespace priors
}  // namespace tamcmc
