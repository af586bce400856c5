/*
 * tamcmc_oracle.h -- CPU oracle for the TAMCMC-C hot path (model spectrum + Whittle logL).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ may be imported, linked or
 * executed by the product (tamcmc-c_b200/, include/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker / the CPU baseline.
 *
 * Every function is a from-scratch plain-C restatement (flat double arrays, no
 * Eigen) of one reference function; the reference file:line it follows is
 * cited next to each definition in tamcmc_oracle.c.  Paths are relative to the
 * upstream repository root (OthmanB/TAMCMC-C v1.86.78).
 *
 * PARITY PINNING: the reference ships no golden vectors for this path and its
 * full build needs Eigen/Boost/GSL (absent here).  The oracle is pinned by
 * compiling the reference's OWN source files for this path (likelihoods.cpp,
 * noise_models.cpp, build_lorentzian.cpp, function_rot.cpp, acoefs.cpp,
 * interpol.cpp, linfit.cpp) against a minimal Eigen-API shim -- see
 * oracle/Makefile target `_ref` and oracle/eigen_shim/ -- and comparing both
 * on seeded inputs (tests/test_oracle_vs_ref.py, tests/golden/).
 */
#ifndef TAMCMC_ORACLE_H
#define TAMCMC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- model ids: Config/default/models_ctrl.list ---- */
enum {
    ORC_MODEL_KALLINGER2014_GAUSSIAN = 0,   /* models.cpp:5728 */
    ORC_MODEL_HARVEY_GAUSSIAN        = 1,   /* models.cpp:5674 */
    ORC_MODEL_MS_GLOBAL_A1L_ETAA3   = 6,
    ORC_MODEL_MS_GLOBAL_A1N_ETAA3   = 7,
    ORC_MODEL_MS_GLOBAL_A1NL_ETAA3  = 8,
    ORC_MODEL_MS_GLOBAL_CLASSIC     = 3,
    ORC_MODEL_MS_LOCAL_BASIC        = 11,
    ORC_MODEL_MS_GLOBAL_CLASSIC_V2  = 12,
    ORC_MODEL_MS_GLOBAL_CLASSIC_V3  = 13,
    ORC_MODEL_MS_LOCAL_HNLM         = 14,
    ORC_MODEL_MS_GLOBAL_A1N_A2A3    = 18,
    ORC_MODEL_MS_GLOBAL_A1NL_A2A3   = 19,
    ORC_MODEL_MS_GLOBAL_AJALM       = 21,
    ORC_MODEL_MS_GLOBAL_AJ          = 23
};

/* status codes: the reference calls exit(); the oracle returns these */
enum {
    ORC_OK = 0,
    ORC_ERR_WINDOW = 1,      /* set_imin_imax: imax-imin<=0 (build_lorentzian.cpp:650-665) */
    ORC_ERR_MODEL = 2,       /* unknown / obsolete model id (model_def.cpp:231-237,352-384) */
    ORC_ERR_ARG = 3
};

/* ---- scalar helpers ---- */
long double orc_Hslm_Ritzoller1991(int s, int l, int m);
long double orc_Pslm(int s, int l, int m);
double orc_Qlm(int l, int m);
int    orc_factorial(int n);
double orc_combi(int n, int r);
double orc_dmm(int l, int m1, int m2, double beta);
void   orc_function_rot(int l, double beta, double *mat /* (2l+1)^2 row-major */);
void   orc_amplitude_ratio(int l, double beta_deg, double *V /* 2l+1 */);
double orc_lin_interpol(const double *x, const double *y, long n, double x_int);
void   orc_linfit(const double *x, const double *y, long n, double out[2]);
double orc_eta0_fct_dnu(double dnu_obs);
double orc_eta0_fct(const double *fl0_all, long n);
void   orc_eval_acoefs(int l, const double *nu_nls, double aj[6]);

/* ---- windows ---- */
int orc_set_imin_imax(const double *x, long N, int l, double fc_l, double gamma_l,
                      double f_s, double c, double step, int ivals[2]);

/* window trace: every optimum_lorentzian_calc_* call appends (l, imin, imax) */
void orc_trace_begin(int *l, int *i0, int *i1, int capacity);
int  orc_trace_end(void); /* returns number of windows recorded */

/* ---- profiles on a window ---- */
void orc_build_l_mode_a1etaa3(const double *x_l, long n, double H_l, double fc_l, double f_s,
                              double eta0, double a3, double asym, double gamma_l, int l,
                              const double *V, double *result);
void orc_build_l_mode_a1etaa3_v2(const double *x_l, long n, const double *H_lm, double fc_l,
                                 double f_s, double eta0, double a3, double asym,
                                 double gamma_l, int l, double *result);
void orc_build_l_mode_a1l_etaa3(const double *x_l, long n, double H_l, double fc_l, double f_s1,
                                double f_s2, double eta0, double a3, double asym,
                                double gamma_l, int l, const double *V, double *result);
void orc_build_l_mode_a1l_a2a3(const double *x_l, long n, double H_l, double fc_l, double f_s1,
                               double f_s2, double a2, double a3, double asym, double gamma_l,
                               int l, const double *V, double *result);
void orc_build_l_mode_aj(const double *x_l, long n, double H_l, double fc_l, double a1, double a2,
                         double a3, double a4, double a5, double a6, double eta0, double asym,
                         double gamma_l, int l, const double *V, double *result);
/* ajAlm with the activity term supplied by the caller: Alm_m[m+l] (GSL is un-vendored) */
void orc_build_l_mode_ajAlm(const double *x_l, long n, double H_l, double fc_l, double a1, double a3,
                            double a5, double eta0, double epsilon_nl, const double *Alm_m,
                            double asym, double gamma_l, int l, const double *V, double *result);

/* ---- windowed accumulation (VectorXd-returning family: y is replaced by a new full-length vector) ---- */
int orc_optimum_lorentzian_calc_a1etaa3(const double *x, long N, double **y_io, double H_l, double fc_l,
                                        double f_s, double eta0, double a3, double asym, double gamma_l,
                                        int l, const double *V, double step, double c);
int orc_optimum_lorentzian_calc_a1etaa3_v2(const double *x, long N, double **y_io, const double *H_lm,
                                           double fc_l, double f_s, double eta0, double a3, double asym,
                                           double gamma_l, int l, double step, double c);
int orc_optimum_lorentzian_calc_a1l_etaa3(const double *x, long N, double **y_io, double H_l, double fc_l,
                                          double f_s1, double f_s2, double eta0, double a3, double asym,
                                          double gamma_l, int l, const double *V, double step, double c);
int orc_optimum_lorentzian_calc_a1l_a2a3(const double *x, long N, double **y_io, double H_l, double fc_l,
                                         double f_s1, double f_s2, double a2, double a3, double asym,
                                         double gamma_l, int l, const double *V, double step, double c);
/* Optim_L family: returns the local block */
typedef struct { double *y; int i0; int N; } orc_Optim_L;
int orc_optimum_lorentzian_calc_aj(const double *x, long N, double H_l, double fc_l, double a1, double a2,
                                   double a3, double a4, double a5, double a6, double eta0, double asym,
                                   double gamma_l, int l, const double *V, double step, double c,
                                   orc_Optim_L *out);

/* ---- noise + likelihood ---- */
void orc_harvey_like(const double *noise_params, int n_noise, const double *x, long N,
                     double **y_io, int Nharvey);
long double orc_likelihood_chi22p(const double *y, const double *model, long N, long p);
long double orc_likelihood_chi_square(const double *y, const double *model, const double *sigma, long N);

/* ---- models (params, plength[11], x[N]) -> model_out[N] ---- */
/* Alm provider for model 21: returns Alm(l,m,theta0,delta) for the selected filter; may be NULL for other models */
typedef double (*orc_alm_fn)(int l, int m, double theta0, double delta, int filter_code, void *user);

int orc_call_model(int model_id, const double *params, const int *plength, const double *x, long N,
                   double *model_out, orc_alm_fn alm, void *alm_user);
/* tempered log-likelihood as Model_def::call_likelihood (model_def.cpp:390-419), likelihood switch 0 */
long double orc_call_likelihood_chi22p(const double *y, const double *model, long N, double p, double Tcoef);

/* one evaluation for Nchains chains (threads over chains like MALA.cpp:648), logL_out[Nchains] */
int orc_eval_chains(int model_id, const double *params, int Nparams, const int *plength,
                    const double *x, const double *y, long N, int Nchains, const double *Tcoefs, double p,
                    double *logL_out, int nthreads);

/* best-effort CPU variant (same arithmetic per element, window-only, single pass, no full-vector copies);
 * used only for the second CPU baseline line in bench.py */
int orc_eval_chains_fast(int model_id, const double *params, int Nparams, const int *plength,
                         const double *x, const double *y, long N, int Nchains, const double *Tcoefs, double p,
                         double *logL_out, int nthreads);

int orc_eval_chains_chi_square(int model_id, const double *params, int Nparams, const int *plength, const double *x,
                               const double *y, const double *sigma, long N, int Nchains, const double *Tcoefs,
                               double *logL_out, int nthreads);

/* generic mode table (TAMCMC_MODEL_MODE_TABLE of include/tamcmc_gpu.h): what the aj-family model functions do after
 * resolving their mode list (models.cpp:4931-5017) */
int orc_mode_table_model(const double *row, int Nnoise, int step_mode, const double *x, long N, double *out);
int orc_mode_table_eval_chains(const double *rows, int row_stride, int Nnoise, int step_mode, const double *x, const double *y,
                               long N, int Nchains, const double *Tcoefs, double p, double *logL_out, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
