// ref_shim_api.cpp -- extern "C" entry points into the REFERENCE's own hot-path functions, compiled where
// they lie under /root/reference against oracle/eigen_shim (see oracle/Makefile, target `ref`).
// TEST INFRASTRUCTURE ONLY: used by tests/test_oracle_vs_reference.py and tests/golden/make_golden_from_reference_cpp.py
// to pin the plain-C oracle.  No reference source is copied here: only declarations from its headers are used.
#include <Eigen/Dense>
#include <cstdlib>
#include <omp.h>
#include <cstring>
#include <string>
#include "build_lorentzian.h"   // reference header: set_imin_imax, build_l_mode_*, optimum_lorentzian_calc_*, Qlm
#include "function_rot.h"       // amplitude_ratio
#include "acoefs.h"             // Pslm, eval_acoefs
#include "interpol.h"           // lin_interpol
#include "linfit.h"             // linfit
#include "noise_models.h"       // harvey_like
#include "likelihoods.h"        // likelihood_chi22p, likelihood_chi_square
#include "models.h"             // model_MS_Global_*, model_MS_local_*, model_RGB_asympt_* (tamcmc/sources/models.cpp)
#include "stats_dictionary.h"   // logP_* primitive priors
#include "priors_calc.h"        // apply_generic_priors, priors_Harvey_Gaussian, priors_Kallinger2014_Gaussian
#include <cmath>
#include "solver_mm.h"            // external/ARMM: solve_mm_asymptotic_O2p, solve_mm_asymptotic_O2from_l0
#include "bump_DP.h"              // external/ARMM: ksi_fct2, h_l_rgb, gamma_l_fct2, dnu_rot_2zones
#include "../../external/spline/src/spline.h"   // tk::spline (resolved against -I$(REFERENCE)/tamcmc/headers)

using Eigen::VectorXd;
using Eigen::VectorXi;

// The Alm activity term: the reference's own external/Alm/Alm_cpp/{activity,Alm_interpol,bilinear_interpol}.cpp are compiled
// in (oracle/Makefile) against stand-ins for the three absent libraries: Boost's spherical harmonics
// (eigen_shim/boost/math/special_functions/spherical_harmonic.hpp), Boost.Iostreams' gzip filter (zlib underneath) and
// GSL's interp2d (eigen_shim/gsl/gsl_interp2d.h: GSL's published bicubic algorithm, NOT GSL itself).
long double Alm(const int, const int, const long double, const long double, std::string);
double Alm_interp_iter_preinitialised(const int, const int, const long double, const long double, const std::string, gsl_funcs);
GridData_Alm_fast loadAllData(const std::string grid_dir, const std::string ftype);
GridData4gsl flatten_grid(const GridData& data);
gsl_interp2d* init_2dgrid(const GridData4gsl& data_flatten);
static external_data g_extra;          // what Config::Config builds once (config.cpp:77-147)
static bool g_extra_ok = false;

// ---- recording wrapper around the reference's optimum_lorentzian_calc_aj (build_lorentzian.cpp:502-522) ----
// build_lorentzian.cpp is compiled with that symbol renamed to refreal_optimum_lorentzian_calc_aj (oracle/Makefile);
// models.cpp therefore calls THIS function, which appends the call's arguments to the active recording (if any) and
// forwards to the reference's implementation.  One row = {l, fc, H, W, a1..a6, eta0, asym, step, c, V[0..6]}.
Optim_L refreal_optimum_lorentzian_calc_aj(const VectorXd& x, const double H_l, const double fc_l, const double a1, const double a2,
                                           const double a3, const double a4, const double a5, const double a6, const double eta0,
                                           const double asym, const double gamma_l, const int l, const VectorXd& V, const double step,
                                           const double c);
static double* g_rec = nullptr;
static int g_rec_cap = 0, g_rec_n = 0;
enum { REC_STRIDE = 21 };
Optim_L optimum_lorentzian_calc_aj(const VectorXd& x, const double H_l, const double fc_l, const double a1, const double a2,
                                   const double a3, const double a4, const double a5, const double a6, const double eta0,
                                   const double asym, const double gamma_l, const int l, const VectorXd& V, const double step,
                                   const double c)
{
    if (g_rec) {
#pragma omp critical(refshim_rec)
        {
            if (g_rec_n < g_rec_cap) {
                double* r = g_rec + (size_t)g_rec_n * REC_STRIDE;
                r[0] = l; r[1] = fc_l; r[2] = H_l; r[3] = gamma_l; r[4] = a1; r[5] = a2; r[6] = a3; r[7] = a4; r[8] = a5; r[9] = a6;
                r[10] = eta0; r[11] = asym; r[12] = step; r[13] = c;
                for (int k = 0; k < 7; k++) r[14 + k] = (k < (int)V.size()) ? V[k] : 0.0;
            }
            g_rec_n++;
        }
    }
    return refreal_optimum_lorentzian_calc_aj(x, H_l, fc_l, a1, a2, a3, a4, a5, a6, eta0, asym, gamma_l, l, V, step, c);
}

static VectorXd vec(const double* p, long n) { VectorXd v(n); for (long i = 0; i < n; i++) v[i] = p[i]; return v; }
static void out(const VectorXd& v, double* p) { for (long i = 0; i < (long)v.size(); i++) p[i] = v[i]; }

extern "C" {

long double ref_Pslm(int s, int l, int m) { return Pslm(s, l, m); }
double ref_Qlm(int l, int m) { return Qlm(l, m); }
void ref_amplitude_ratio(int l, double beta, double* V) { out(amplitude_ratio(l, beta), V); }
double ref_lin_interpol(const double* x, const double* y, long n, double xi) { return lin_interpol(vec(x, n), vec(y, n), xi); }
void ref_linfit(const double* x, const double* y, long n, double* o) { out(linfit(vec(x, n), vec(y, n)), o); }
void ref_eval_acoefs(int l, const double* nu, double* aj) { VectorXd v = vec(nu, 2 * l + 1); out(eval_acoefs(l, v), aj); }

void ref_set_imin_imax(const double* x, long N, int l, double fc, double gamma, double f_s, double c, double step, int* iv)
{
    // NB: the reference exits the process when imax - imin <= 0; callers only pass valid windows
    VectorXi r = set_imin_imax(vec(x, N), l, fc, gamma, f_s, c, step);
    iv[0] = r[0]; iv[1] = r[1];
}

void ref_build_l_mode_a1etaa3(const double* xl, long n, double H, double fc, double f_s, double eta0, double a3, double asym,
                              double gamma, int l, const double* V, double* res)
{ out(build_l_mode_a1etaa3(vec(xl, n), H, fc, f_s, eta0, a3, asym, gamma, l, vec(V, 2 * l + 1)), res); }

void ref_build_l_mode_a1etaa3_v2(const double* xl, long n, const double* Hlm, double fc, double f_s, double eta0, double a3,
                                 double asym, double gamma, int l, double* res)
{ out(build_l_mode_a1etaa3_v2(vec(xl, n), vec(Hlm, 2 * l + 1), fc, f_s, eta0, a3, asym, gamma, l), res); }

void ref_build_l_mode_a1l_etaa3(const double* xl, long n, double H, double fc, double f_s1, double f_s2, double eta0, double a3,
                                double asym, double gamma, int l, const double* V, double* res)
{ out(build_l_mode_a1l_etaa3(vec(xl, n), H, fc, f_s1, f_s2, eta0, a3, asym, gamma, l, vec(V, 2 * l + 1)), res); }

void ref_build_l_mode_a1l_a2a3(const double* xl, long n, double H, double fc, double f_s1, double f_s2, double a2, double a3,
                               double asym, double gamma, int l, const double* V, double* res)
{ out(build_l_mode_a1l_a2a3(vec(xl, n), H, fc, f_s1, f_s2, a2, a3, asym, gamma, l, vec(V, 2 * l + 1)), res); }

void ref_build_l_mode_aj(const double* xl, long n, double H, double fc, double a1, double a2, double a3, double a4, double a5,
                         double a6, double eta0, double asym, double gamma, int l, const double* V, double* res)
{ out(build_l_mode_aj(vec(xl, n), H, fc, a1, a2, a3, a4, a5, a6, eta0, asym, gamma, l, vec(V, 2 * l + 1)), res); }

// y_out = optimum_lorentzian_calc_a1etaa3(x, y, ...): the full-length vector the reference returns
void ref_optimum_lorentzian_calc_a1etaa3(const double* x, const double* y, long N, double H, double fc, double f_s, double eta0,
                                         double a3, double asym, double gamma, int l, const double* V, double step, double c, double* y_out)
{ out(optimum_lorentzian_calc_a1etaa3(vec(x, N), vec(y, N), H, fc, f_s, eta0, a3, asym, gamma, l, vec(V, 2 * l + 1), step, c), y_out); }

// Optim_L variant: returns i0 and N, fills block[N]
void ref_optimum_lorentzian_calc_aj(const double* x, long N, double H, double fc, double a1, double a2, double a3, double a4,
                                    double a5, double a6, double eta0, double asym, double gamma, int l, const double* V,
                                    double step, double c, int* i0, int* n, double* block)
{
    Optim_L r = optimum_lorentzian_calc_aj(vec(x, N), H, fc, a1, a2, a3, a4, a5, a6, eta0, asym, gamma, l, vec(V, 2 * l + 1), step, c);
    *i0 = r.i0; *n = r.N; out(r.y, block);
}

void ref_harvey_like(const double* noise, int n_noise, const double* x, const double* y, long N, int Nharvey, double* y_out)
{ out(harvey_like(vec(noise, n_noise), vec(x, N), vec(y, N), Nharvey), y_out); }

long double ref_likelihood_chi22p(const double* y, const double* model, long N, long p) { return likelihood_chi22p(vec(y, N), vec(model, N), p); }
long double ref_likelihood_chi_square(const double* y, const double* model, const double* sigma, long N)
{ return likelihood_chi_square(vec(y, N), vec(model, N), vec(sigma, N)); }

// The reference's own model functions (tamcmc/sources/models.cpp), selected like Model_def::call_model
// (tamcmc/sources/model_def.cpp:220-388).  Returns 0, or 2 for ids that are obsolete/unknown there or need the
// GSL-backed Alm grids (21).
int ref_call_model(int model_id, const double* params, int nparams, const int* plength, const double* x, long N, double* model_out)
{
    VectorXd p = vec(params, nparams), xv = vec(x, N), m;
    VectorXi pl(11);
    for (int i = 0; i < 11; i++) pl[i] = plength[i];
    switch (model_id) {
    case 0: m = model_Kallinger2014_Gaussian(p, pl, xv, false); break;      // NB: rewrites ./params.model on every call (models.cpp:5764)
    case 1: m = model_Harvey_Gaussian(p, pl, xv, false); break;
    case 3: m = model_MS_Global_a1etaa3_HarveyLike_Classic(p, pl, xv, false); break;
    case 6: m = model_MS_Global_a1l_etaa3_HarveyLike(p, pl, xv, false); break;
    case 7: m = model_MS_Global_a1n_etaa3_HarveyLike(p, pl, xv, false); break;
    case 8: m = model_MS_Global_a1nl_etaa3_HarveyLike(p, pl, xv, false); break;
    case 11: m = model_MS_local_basic(p, pl, xv, false); break;
    case 12: m = model_MS_Global_a1etaa3_HarveyLike_Classic_v2(p, pl, xv, false); break;
    case 13: m = model_MS_Global_a1etaa3_HarveyLike_Classic_v3(p, pl, xv, false); break;
    case 14: m = model_MS_local_Hnlm(p, pl, xv, false); break;
    case 18: m = model_MS_Global_a1n_a2a3_HarveyLike(p, pl, xv, false); break;
    case 19: m = model_MS_Global_a1nl_a2a3_HarveyLike(p, pl, xv, false); break;
    case 21:                                                               // needs ref_alm_grids_load() first (model_def.cpp:322-324)
        if (!g_extra_ok) return 3;
        m = model_MS_Global_ajAlm_HarveyLike(p, pl, xv, false, g_extra);
        break;
    case 23: m = model_MS_Global_aj_HarveyLike(p, pl, xv, false); break;
    case 25: m = model_RGB_asympt_aj_AppWidth_HarveyLike_v4(p, pl, xv, false); break;
    case 27: m = model_RGB_asympt_aj_CteWidth_HarveyLike_v4(p, pl, xv, false); break;
    default: return 2;
    }
    out(m, model_out);
    return 0;
}

// ---- the ARMM mixed-mode solver and its helpers (external/ARMM), as the red-giant models call them ----
static int put(const VectorXd& v, double* dst, int cap) { if ((int)v.size() > cap) return -1; for (long i = 0; i < (long)v.size(); i++) dst[i] = v[i]; return (int)v.size(); }
int ref_solve_mm_from_l0(const double* nu_l0, int n, int el, double delta0l, double DPl, double alpha, double q, double resol, double fmin, double fmax,
                         int cap, double* nu_m, int* n_m, double* nu_p, double* dnup, int* n_p, double* nu_g, int* n_g)
{
    Data_eigensols r = solve_mm_asymptotic_O2from_l0(vec(nu_l0, n), el, delta0l, DPl, alpha, q, 0, resol, true, false, fmin, fmax);
    *n_m = put(r.nu_m, nu_m, cap); *n_p = put(r.nu_p, nu_p, cap); put(r.dnup, dnup, cap); *n_g = put(r.nu_g, nu_g, cap);
    return (*n_m < 0 || *n_p < 0 || *n_g < 0) ? 1 : 0;
}
int ref_solve_mm_O2p(double Dnu_p, double epsilon, int el, double delta0l, double alpha_p, double nmax, double DPl, double alpha, double q, double fmin,
                     double fmax, double resol, int cap, double* nu_m, int* n_m, double* nu_p, double* dnup, int* n_p, double* nu_g, int* n_g)
{
    Data_eigensols r = solve_mm_asymptotic_O2p(Dnu_p, epsilon, el, delta0l, alpha_p, nmax, DPl, alpha, q, 0, fmin, fmax, resol, true, false);
    *n_m = put(r.nu_m, nu_m, cap); *n_p = put(r.nu_p, nu_p, cap); put(r.dnup, dnup, cap); *n_g = put(r.nu_g, nu_g, cap);
    return (*n_m < 0 || *n_p < 0 || *n_g < 0) ? 1 : 0;
}
void ref_ksi_fct2(const double* nu, int n, const double* nu_p, const double* dnup, int n_p, const double* nu_g, const double* dPg, int n_g, double q, double* ksi)
{ out(ksi_fct2(vec(nu, n), vec(nu_p, n_p), vec(nu_g, n_g), vec(dnup, n_p), vec(dPg, n_g), q, "precise"), ksi); }
// tk::spline exactly as the models set it up (models.cpp:4834-4843): type 1 cspline, 2 cspline_hermite
void ref_spline_eval(const double* x, const double* y, int n, int type, const double* xq, int nq, double* o)
{
    tk::spline s;
    s.set_boundary(tk::spline::second_deriv, 0.0, tk::spline::second_deriv, 0.0);
    s.set_points(std::vector<double>(x, x + n), std::vector<double>(y, y + n), type == 1 ? tk::spline::cspline : tk::spline::cspline_hermite);
    for (int i = 0; i < nq; i++) o[i] = s(xq[i]);
}

// ---- the Alm activity term through the reference's own sources ----
static const char* ftype_of(int filter_code) { return filter_code == 0 ? "gate" : filter_code == 1 ? "gauss" : "triangle"; }

// Alm() direct integral, activity.cpp:221-246 (theta0, delta in radians)
double ref_Alm(int l, int m, double theta0, double delta, int filter_code) { return (double)Alm(l, m, theta0, delta, ftype_of(filter_code)); }

// the grid set-up of Config::Config (config.cpp:77-147) with the reference's own loadAllData / flatten_grid / init_2dgrid
static bool fill_funcs(const std::string& dir, const char* ftype, gsl_funcs& f)
{
    GridData_Alm_fast g = loadAllData(dir, ftype);
    if (g.error) return false;
    f.flat_grid_A10 = flatten_grid(g.A10); f.flat_grid_A11 = flatten_grid(g.A11);
    f.flat_grid_A20 = flatten_grid(g.A20); f.flat_grid_A21 = flatten_grid(g.A21); f.flat_grid_A22 = flatten_grid(g.A22);
    f.flat_grid_A30 = flatten_grid(g.A30); f.flat_grid_A31 = flatten_grid(g.A31); f.flat_grid_A32 = flatten_grid(g.A32);
    f.flat_grid_A33 = flatten_grid(g.A33);
    f.interp_A10 = init_2dgrid(f.flat_grid_A10); f.interp_A11 = init_2dgrid(f.flat_grid_A11);
    f.interp_A20 = init_2dgrid(f.flat_grid_A20); f.interp_A21 = init_2dgrid(f.flat_grid_A21); f.interp_A22 = init_2dgrid(f.flat_grid_A22);
    f.interp_A30 = init_2dgrid(f.flat_grid_A30); f.interp_A31 = init_2dgrid(f.flat_grid_A31); f.interp_A32 = init_2dgrid(f.flat_grid_A32);
    f.interp_A33 = init_2dgrid(f.flat_grid_A33);
    f.valid = true;
    return true;
}
int ref_alm_grids_load(const char* grid_dir)
{
    g_extra_ok = fill_funcs(grid_dir, "gate", g_extra.Alm_interp_gate) && fill_funcs(grid_dir, "triangle", g_extra.Alm_interp_triangle);
    return g_extra_ok ? 0 : 1;
}
// Alm_interp_iter_preinitialised, Alm_interpol.cpp:188-348
double ref_Alm_interp(int l, int m, double theta0, double delta, int filter_code)
{
    if (!g_extra_ok) return std::nan("");
    return Alm_interp_iter_preinitialised(l, m, theta0, delta, ftype_of(filter_code),
                                          filter_code == 0 ? g_extra.Alm_interp_gate : g_extra.Alm_interp_triangle);
}

// One MCMC step's worth of likelihood work with the reference's own functions: the per-chain OpenMP fan-out of
// MALA.cpp:648 around call_model + call_likelihood (model_def.cpp:466-482, 390-401): model.row(m) = model_X(...);
// logL = likelihood_chi22p(y, model.row(m), p) / Tcoefs[m].  `model` is kept as the reference's MatrixXd(Nchains, N)
// so the strided row store / row read are part of the timed work like in the reference.
int ref_eval_chains(int model_id, const double* params, int nparams, const int* plength, const double* x, const double* y,
                    long N, int Nchains, const double* Tcoefs, double p_like, double* logL_out, int nthreads)
{
    const VectorXd xv = vec(x, N), yv = vec(y, N);
    VectorXi pl(11);
    for (int i = 0; i < 11; i++) pl[i] = plength[i];
    Eigen::MatrixXd model(Nchains, N);
    int rc = 0;
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for default(shared)
    for (int chain = 0; chain < Nchains; chain++) {
        const VectorXd pv = vec(params + (size_t)chain * nparams, nparams);
        VectorXd m;
        switch (model_id) {
        case 3: m = model_MS_Global_a1etaa3_HarveyLike_Classic(pv, pl, xv, false); break;
        case 6: m = model_MS_Global_a1l_etaa3_HarveyLike(pv, pl, xv, false); break;
        case 11: m = model_MS_local_basic(pv, pl, xv, false); break;
        case 12: m = model_MS_Global_a1etaa3_HarveyLike_Classic_v2(pv, pl, xv, false); break;
        case 13: m = model_MS_Global_a1etaa3_HarveyLike_Classic_v3(pv, pl, xv, false); break;
        case 23: m = model_MS_Global_aj_HarveyLike(pv, pl, xv, false); break;
        case 25: m = model_RGB_asympt_aj_AppWidth_HarveyLike_v4(pv, pl, xv, false); break;
        case 27: m = model_RGB_asympt_aj_CteWidth_HarveyLike_v4(pv, pl, xv, false); break;
        default: rc = 2; continue;
        }
        model.row(chain) = m;
        const double p = p_like;
        const long double logL = likelihood_chi22p(yv, model.row(chain), p);
        logL_out[chain] = (double)(logL / Tcoefs[chain]);
    }
    return rc;
}

int ref_max_threads(void) { return omp_get_max_threads(); }

// ref_call_model while recording every optimum_lorentzian_calc_aj call the model function makes: rows[cap][21]
// (see REC_STRIDE above); *nrows receives the number of calls.  Not thread-safe (one recording at a time).
int ref_call_model_recorded(int model_id, const double* params, int nparams, const int* plength, const double* x, long N,
                            double* model_out, double* rows, int cap, int* nrows)
{
    g_rec = rows; g_rec_cap = cap; g_rec_n = 0;
    const int rc = ref_call_model(model_id, params, nparams, plength, x, N, model_out);
    *nrows = g_rec_n;
    g_rec = nullptr; g_rec_cap = 0;
    return rc;
}

double ref_eta0_fct(const double* fl0, long n) { return eta0_fct(vec(fl0, n)); }

// primitive priors by their switch value (Config/default/primepriors_ctrl.list), stats_dictionary.cpp
double ref_logP(int kind, double a, double b, double c, double d, double x)
{
    switch (kind) {
    case 1: return (double)logP_uniform(a, b, x);
    case 2: return (double)logP_gaussian(a, b, x);
    case 4: return (double)logP_jeffrey(a, b, x);
    case 5: return (double)logP_uniform_gaussian(a, b, c, x);
    case 6: return (double)logP_gaussian_uniform(a, b, c, x);
    case 7: return (double)logP_gaussian_uniform_gaussian(a, b, c, d, x);
    case 8: return (double)logP_uniform_abs(a, b, x);
    case 9: return (double)logP_uniform_cos(a, b, x);
    case 10: return (double)logP_jeffrey_abs(a, b, x);
    }
    return 0.0;
}

// 1-D tabulated prior (switch value 11), stats_dictionary.cpp:252-291
double ref_logP_tabulated(const double* tab_x, const double* tab_y, int n, double x, int normalise)
{
    return (double)logP_tabulated(vec(tab_x, n), vec(tab_y, n), x, normalise != 0);
}

// which = -1: apply_generic_priors; 0 / 1: priors_Kallinger2014_Gaussian / priors_Harvey_Gaussian (priors_calc.cpp:725, 649, 631).
// pri = [4][n] row-major (the reference's MatrixXd priors_params(4, n)), kinds = priors_names_switch.
double ref_priors(int which, const double* params, int n, const double* pri, const int* kinds)
{
    VectorXd p = vec(params, n);
    MatrixXd P(4, n);
    VectorXi k(n), pl(11);
    for (int i = 0; i < n; i++) { k[i] = kinds[i]; for (int r = 0; r < 4; r++) P(r, i) = pri[(size_t)r * n + i]; }
    for (int i = 0; i < 11; i++) pl[i] = 0;
    tabpriors none;
    if (which == 0) return (double)priors_Kallinger2014_Gaussian(p, pl, P, k, none);
    if (which == 1) return (double)priors_Harvey_Gaussian(p, pl, P, k, none);
    return (double)apply_generic_priors(p, P, k, none);
}

}  // extern "C"
