/*
 * tamcmc_gpu.h -- C ABI of the B200-native hot path of TAMCMC-C:
 * power-spectrum model + Whittle chi^2(2 dof) log-likelihood for every parallel-tempered chain.
 *
 * The reference (OthmanB/TAMCMC-C v1.86.78) has no FFI for this path; its "operator API" is the
 * C++ switch Model_def::call_model / call_likelihood over free functions.  Each entry point below
 * names the reference interface it replaces (file:line relative to the upstream repo root).
 * INTEGRATION.md shows the binding a maintainer adds in tamcmc/sources/model_def.cpp.
 *
 * Conventions: plain pointers and sizes, caller-owned output buffers, no allocation on the
 * evaluation path, never exit(): every call returns a tamcmc_status.  A context is not
 * re-entrant; one host thread issues one batched call per MCMC step.
 * There is NO CPU fallback: with no usable CUDA device every call fails with TAMCMC_ERR_CUDA.
 */
#ifndef TAMCMC_GPU_H
#define TAMCMC_GPU_H

#ifdef __cplusplus
extern "C" {
#endif

#define TAMCMC_GPU_ABI_VERSION 2

typedef enum {
    TAMCMC_OK = 0,
    TAMCMC_ERR_ARG = 1,          /* bad pointer / size / plength inconsistent with Nparams */
    TAMCMC_ERR_MODEL = 2,        /* model id unknown, obsolete or not on this path
                                    (reference: exit(), model_def.cpp:231-237, 352-384) */
    TAMCMC_ERR_CUDA = 3,         /* CUDA runtime error, or no device: tamcmc_gpu_last_error() has the text */
    TAMCMC_ERR_WINDOW = 4,       /* some chain hit imax-imin<=0 in set_imin_imax
                                    (reference: exit(), build_lorentzian.cpp:650-665); its logL is NaN,
                                    status_out[] tells which chain */
    TAMCMC_ERR_NONFINITE = 5,    /* some chain produced a non-finite mode quantity; its logL is NaN */
    TAMCMC_ERR_LIKELIHOOD = 6,   /* likelihood id not on this path (model_def.cpp:405-416) */
    TAMCMC_ERR_POOL = 7          /* reserved (ABI v1 reported an exhausted device list pool; tile lists now live in shared
                                    memory only and the code is never returned) */
} tamcmc_status;

/* per-chain status bits written to status_out[] (0 = evaluated normally) */
#define TAMCMC_CHAIN_WINDOW     1
#define TAMCMC_CHAIN_NONFINITE  2
#define TAMCMC_CHAIN_BADCFG     4
#define TAMCMC_CHAIN_INACTIVE   8   /* active_mask[chain]==0 */

/* model ids = case labels of Model_def::call_model (model_def.cpp:220-388) =
 * Config/default/models_ctrl.list */
/* Gaussian-envelope models (no Lorentzians; plength is not read, parameters sit at fixed positions):
 *   0: [k_a, s_a, k_b0, s_b0, c0, a1, a2, k1, s1, c1, k2, s2, c2, N0, Amax, numax, sigma, mu_numax] (Nparams >= 18); the
 *      super-Lorentzian normalisations ksi_k integrate over the whole spectrum (noise_models.cpp:65-84): no bin-range slices
 *   1: [H1, tc1, p1, H2, tc2, p2, B0, Hgauss, nu_gauss, sigma] (Nparams >= 10) */
#define TAMCMC_MODEL_KALLINGER2014_GAUSSIAN                   0   /* models.cpp:5728 */
#define TAMCMC_MODEL_HARVEY_GAUSSIAN                          1   /* models.cpp:5674 */
#define TAMCMC_MODEL_MS_GLOBAL_A1ETAA3_HARVEYLIKE_CLASSIC     3   /* models.cpp:1943 */
#define TAMCMC_MODEL_MS_GLOBAL_A1L_ETAA3_HARVEYLIKE           6   /* models.cpp:25   */
#define TAMCMC_MODEL_MS_GLOBAL_A1N_ETAA3_HARVEYLIKE            7   /* models.cpp:217  */
#define TAMCMC_MODEL_MS_GLOBAL_A1NL_ETAA3_HARVEYLIKE           8   /* models.cpp:1003 */
#define TAMCMC_MODEL_MS_LOCAL_BASIC                          11   /* models.cpp:3012 */
#define TAMCMC_MODEL_MS_GLOBAL_A1ETAA3_HARVEYLIKE_CLASSIC_V2 12   /* models.cpp:2128 */
#define TAMCMC_MODEL_MS_GLOBAL_A1ETAA3_HARVEYLIKE_CLASSIC_V3 13   /* models.cpp:2338 */
#define TAMCMC_MODEL_MS_LOCAL_HNLM                           14   /* models.cpp:3198 */
/* ids 18 / 19 (model_MS_Global_a1n_a2a3 / a1nl_a2a3_HarveyLike) print "not tested yet" and exit in the reference
 * (models.cpp:599-603, 993-997): TAMCMC_ERR_MODEL */
#define TAMCMC_MODEL_MS_GLOBAL_AJ_HARVEYLIKE                 23   /* models.cpp:1195 */

/* Generic MODE TABLE (not a reference model id): the entry for model functions whose mode list is resolved by host
 * code that stays in the reference -- the ARMM mixed-mode solver of model_RGB_asympt_aj_*_HarveyLike_v4
 * (models.cpp:4684-5079), the GSL Alm grids of model_MS_Global_ajAlm_HarveyLike (models.cpp:1411-1746).  The host passes
 * exactly what those functions pass to optimum_lorentzian_calc_aj / _ajAlm (build_lorentzian.cpp:502-522, 480-499) and to
 * harvey_like (noise_models.cpp:15-39); windows, profiles, background and likelihood run on the GPU.
 *   plength = [capacity = max modes per chain, step_mode, 0,0,0,0,0,0, Nnoise, 0, 0]
 *             step_mode 0: step = x[1]-x[0] (MS models, models.cpp:1952); 1: step = x[2]-x[1] (RGB v4, models.cpp:4714)
 *   Nparams = TAMCMC_MT_HEADER + Nnoise + TAMCMC_MT_STRIDE * capacity
 *   row     = [nmodes, inclination(deg), trunc_c, asym, noise[Nnoise] (abs() is applied like models.cpp:5017),
 *              nmodes x {l, fc, H_l, gamma_l, a1, a2, a3, a4, a5, a6, eta0, extra[m=-3..3], 0, 0}]
 *             m-heights are H_l * amplitude_ratio(l, inclination)[m]; the window uses a1 as f_s;
 *             nu_nlm = build_l_mode_aj's (build_lorentzian.cpp:222-226) + extra[m]  (extra = fc*epsilon_nl*Alm(l,m) reproduces
 *             build_l_mode_ajAlm, build_lorentzian.cpp:182-190).  H_l and gamma_l must be >= 0 (the model functions pass
 *             abs() values); the same per-mode columns as the reference's own `mode_params` table (models.cpp:4941). */
#define TAMCMC_MODEL_MODE_TABLE 1000
#define TAMCMC_MT_HEADER 4
#define TAMCMC_MT_STRIDE 20

/* likelihood ids = Config/default/likelihoods_ctrl.list (model_def.cpp:396-403) */
#define TAMCMC_LIKELIHOOD_CHI22P 0      /* likelihood_chi22p, likelihoods.cpp:17-28 */
#define TAMCMC_LIKELIHOOD_CHI_SQUARE 1  /* likelihood_chi_square, likelihoods.cpp:31-40: -sum((y-M)^2/sigma_y^2)/2, /Tcoefs[m] */

typedef struct tamcmc_gpu_ctx tamcmc_gpu_ctx;

/* One star (or one slice of a star) = the reference's `Data` {x, y, Nx} (headers/data.h) plus the
 * model selection of its Model_def (model_fct_name_switch, plength: model_def.h:40-66). */
typedef struct {
    int model_id;
    int plength[11];        /* io_ms_global.cpp:1315-1325 */
    int Nparams;            /* length of one parameter row */
    const double *x;        /* host, length N: frequencies (microHz), regular grid */
    const double *y;        /* host, length N: power spectrum */
    long N;                 /* bins held by THIS context */
    /* bin-sharding over GPUs (leave zero for a whole spectrum): this context holds global bins
     * [bin_offset, bin_offset+N) of a spectrum of N_global bins whose first two and last
     * frequencies are x_first, x_second, x_last (needed by set_imin_imax and `step`). */
    long N_global;
    long bin_offset;
    double x_first, x_second, x_last;
    /* chi_square likelihood only: uncertainties of y, length N (Data.sigma_y); NULL = all ones, which is what the
     * reference substitutes when the data file has no such column (config.cpp:367-374) */
    const double *sigma_y;
} tamcmc_gpu_star;

/* Replaces: the data/model set-up of Model_def::Model_def (model_def.cpp:28-160) for the hot path.
 * Uploads x, y once; allocates every device buffer the evaluation path needs.
 * Tcoefs[Nchains]: tempering coefficients (MALA.cpp:75-80); p: likelihood_params (model_def.cpp:399). */
int tamcmc_gpu_create(int device, int nstars, const tamcmc_gpu_star *stars, int Nchains,
                      const double *Tcoefs, double p, int likelihood_id, tamcmc_gpu_ctx **out);

void tamcmc_gpu_destroy(tamcmc_gpu_ctx *ctx);

/* Replaces: the per-chain OpenMP fan-out of call_model + call_likelihood inside
 * Model_def::generate_model (model_def.cpp:466-482) driven by MALA::update_position_MH
 * (MALA.cpp:648-668): ONE batched evaluation of all chains of all stars.
 *   params      host [nstars][Nchains][params_stride] row-major, params_stride = tamcmc_gpu_params_stride()
 *   active_mask host [nstars][Nchains] or NULL; 0 reproduces the logPrior==-inf short-circuit
 *               (model_def.cpp:476-480): the chain is skipped and logL_out is NaN
 *   logL_out    host [nstars][Nchains]: TEMPERED log-likelihood -p*S/Tcoefs[m] (model_def.cpp:401)
 *   status_out  host [nstars][Nchains] or NULL: TAMCMC_CHAIN_* bits
 * Host<->device copies of params and results are part of this call. */
int tamcmc_gpu_eval(tamcmc_gpu_ctx *ctx, const double *params, const unsigned char *active_mask,
                    double *logL_out, int *status_out);

/* The same evaluation in two halves, for callers with host work that does not depend on the result (e.g. drawing the next
 * proposal's random numbers while the GPU evaluates this one): _begin stages `params` / `active_mask` (the caller's buffers are
 * free again on return) and launches; _end waits, fills logL_out / status_out and returns what tamcmc_gpu_eval returns.
 * tamcmc_gpu_eval == _begin + _end.  One evaluation in flight per context: a second _begin before _end is TAMCMC_ERR_ARG. */
int tamcmc_gpu_eval_begin(tamcmc_gpu_ctx *ctx, const double *params, const unsigned char *active_mask);

/* The context's own staging block for parameter rows: pinned host memory, [nstars][Nchains][*stride_out] doubles, valid
 * until tamcmc_gpu_destroy.  A caller that BUILDS its rows (the host expanders tamcmc_host_expand_rgb_v4 /
 * tamcmc_host_expand_ajAlm, or Model_def-style code that assembles `params` per chain, model_def.cpp:466-482) writes them here
 * and passes this pointer as `params` to tamcmc_gpu_eval / _eval_begin: the call then skips its host-side copy of the block
 * (126 KB per step for ten 78-mode red-giant mode tables).  The block must not be rewritten between _begin and _end. */
double *tamcmc_gpu_params_staging(tamcmc_gpu_ctx *ctx, int *stride_out);
int tamcmc_gpu_eval_end(tamcmc_gpu_ctx *ctx, double *logL_out, int *status_out);

/* Same evaluation with DEVICE-resident inputs/outputs on the caller's CUDA stream (cudaStream_t
 * passed as void*; NULL = the context's own stream).  d_params has the layout of `params` above,
 * d_logL is [nstars][Nchains].  Asynchronous: the caller synchronises the stream.
 * If raw_sum != 0, d_logL receives S = sum_i(ln M_i + y_i/M_i) over the bins held by this context
 * (the quantity a bin-sharded run all-reduces before applying -p/T). */
int tamcmc_gpu_eval_device(tamcmc_gpu_ctx *ctx, const double *d_params, const unsigned char *d_active,
                           double *d_logL, int raw_sum, void *stream);

/* Replaces: MALA::parallel_tempering (MALA.cpp:397-461) for callers that keep chains on the device (e.g. a bin-sharded run
 * whose ranks all hold the all-reduced log-likelihoods): swap of the adjacent chains A and A+1 of star `star`, decided and
 * applied on the device, asynchronously on `stream` (NULL = the context's stream).  d_logL holds TEMPERED log-likelihoods
 * (layout of tamcmc_gpu_eval_device): with LA' = logL[A] T[A]/T[A+1] and LB' = logL[A+1] T[A+1]/T[A], the swap is accepted when
 * u <= min(1, exp(LA' + LB' - logL[A] - logL[A+1])) (NaN never accepts); then the parameter rows A and A+1 of d_params and, if
 * given, the entries of d_logPrior are exchanged and logL[A] = LB', logL[A+1] = LA'.  `u` is the caller's uniform draw in [0, 1)
 * (every rank of a sharded run passes the same one); d_swapped (device int, may be NULL) receives 1 or 0. */
int tamcmc_gpu_pt_swap_device(tamcmc_gpu_ctx *ctx, int star, int A, double u, double *d_params, double *d_logL,
                              double *d_logPrior, int *d_swapped, void *stream);

/* ---- the exchange step of a bin-sharded spectrum (one sum per chain, <= 192 bytes) over NVLink peer memory ----
 * Every rank holds a context created with tamcmc_gpu_star.N_global / bin_offset (its bin range of the same spectrum, same
 * Nchains, Tcoefs, p, likelihood).  Once the exchange is attached, an evaluation with raw_sum == 0 returns the log-likelihood
 * of the WHOLE spectrum on every rank: the last CTA of the rank's fused kernel writes its chains' local sums S into the exchange
 * buffer of every rank (peer stores), publishes a flag, waits for the other ranks' flags, adds the sums in rank order (bitwise
 * identical results on all ranks, run to run) and applies -p S / Tcoefs[m] (model_def.cpp:399-401) -- no collective launch, no
 * host round trip.  All ranks must evaluate in lock step (same number of evaluations); a rank that never arrives ends the wait
 * after ~2 s with NaN results and TAMCMC_ERR_NONFINITE.
 *   _create : allocates this rank's exchange buffer; handle_out receives its 64-byte CUDA IPC handle, to be sent to the other
 *             ranks by whatever the caller uses for plumbing (torch.distributed all_gather, MPI, a file)
 *   _attach : handles = world x 64 bytes, rank order (entry `rank` is ignored); opens the peers' buffers (cudaIpcOpenMemHandle)
 *   _attach_ptrs : the same for ranks that live in ONE process (e.g. one thread per GPU): bufs[r] = tamcmc_gpu_exchange_buffer()
 *             of rank r's context, peer access already enabled by the caller */
#define TAMCMC_XCHG_HANDLE_BYTES 64
#define TAMCMC_XCHG_MAX_RANKS 8
int tamcmc_gpu_exchange_create(tamcmc_gpu_ctx *ctx, void *handle_out);
int tamcmc_gpu_exchange_attach(tamcmc_gpu_ctx *ctx, int rank, int world, const void *handles);
int tamcmc_gpu_exchange_attach_ptrs(tamcmc_gpu_ctx *ctx, int rank, int world, void *const *bufs);
void *tamcmc_gpu_exchange_buffer(tamcmc_gpu_ctx *ctx);

/* Waits for the context's own stream (after tamcmc_gpu_eval_device with stream == NULL) and, when
 * profiling is on, accumulates the CUDA-event durations of that evaluation's kernels. */
int tamcmc_gpu_sync(tamcmc_gpu_ctx *ctx);

/* Replaces: Model_def::call_model_explicit (model_def.cpp:209-218) as used by tools/getmodel.cpp:201
 * and the diagnostics mean-model (MALA.cpp:722,735): the model spectrum for ONE parameter row.
 * model_out: host, N doubles of star `star`. */
int tamcmc_gpu_model(tamcmc_gpu_ctx *ctx, int star, const double *params_row, double *model_out);

/* Parity/debug: the bin windows of set_imin_imax (build_lorentzian.cpp:595-676) for one parameter
 * row, in the reference's call order.  Arrays of capacity `cap`; *nmodes receives the count. */
int tamcmc_gpu_windows(tamcmc_gpu_ctx *ctx, int star, const double *params_row, int cap,
                       int *nmodes, int *l, int *imin, int *imax);

/* Parity/debug: the expanded component table (nu_nlm, height H*V_m, width) for one parameter row. */
int tamcmc_gpu_components(tamcmc_gpu_ctx *ctx, int star, const double *params_row, int cap,
                          int *ncomp, int *mode_index, int *m, double *nu, double *height, double *width);

/* sizes */
int  tamcmc_gpu_params_stride(const tamcmc_gpu_ctx *ctx);   /* doubles per parameter row (max Nparams) */
int  tamcmc_gpu_nstars(const tamcmc_gpu_ctx *ctx);
int  tamcmc_gpu_nchains(const tamcmc_gpu_ctx *ctx);
long tamcmc_gpu_pairs_last(tamcmc_gpu_ctx *ctx);            /* sum over chains and modes of (2l+1)*(imax-imin)
                                                               for the last evaluation (SURVEY.md 8d "P") */

/* measurement support: CUDA-event timing of the kernels on the context's stream */
int tamcmc_gpu_set_profiling(tamcmc_gpu_ctx *ctx, int on);
int tamcmc_gpu_get_kernel_ms(tamcmc_gpu_ctx *ctx, long *nlaunch, double *expand_ms_total, double *whittle_ms_total);
long tamcmc_gpu_launch_count(const tamcmc_gpu_ctx *ctx);    /* kernels launched by this context so far */
/* profiling aid, only in libraries built with -DTAMCMC_TRACE (TAMCMC_ERR_ARG otherwise): per-CTA %globaltimer
 * stamps of the last fused-kernel launch, 64 slots per CTA: [0] start, [1] end, [2+2i]/[3+2i] begin/end of the
 * i-th wait for a work segment */
int tamcmc_gpu_debug_trace(tamcmc_gpu_ctx *ctx, unsigned long long *out, int nctas);
/* DFMA microbenchmark: achieved FP64 TFLOP/s of `device` (FMA = 2 flops); the roofline denominator */
int tamcmc_gpu_fp64_peak(int device, double *tflops);

/* ---- host expanders (plain host code inside the same library; they produce MODE TABLE rows) ---- */
/* Alm(l, m, theta0, delta) in radians for filter_code 0 ("gate") / 2 ("triangle"); user callbacks have this signature */
typedef double (*tamcmc_alm_fn)(int l, int m, double theta0, double delta, int filter_code, void *user);
/* Replaces: Alm() of external/Alm/Alm_cpp/activity.cpp:221-246 (direct integral; same 64-point Gauss-Legendre rule in theta) */
double tamcmc_host_alm(int l, int m, double theta0, double delta, int filter_code);
/* Replaces: the host half of model_MS_Global_ajAlm_HarveyLike (models.cpp:1411-1746; model id 21 of models_ctrl.list):
 * params/plength in that model's layout -> one MODE TABLE row of `capacity` modes (row_out must hold
 * TAMCMC_MT_HEADER + Nnoise + TAMCMC_MT_STRIDE*capacity doubles).  alm == NULL uses tamcmc_host_alm; pass a callback to
 * use the reference's GSL grid interpolation instead (Alm_interp_iter_preinitialised, bilinear_interpol.cpp:118-130). */
int tamcmc_host_expand_ajAlm(const double *params, const int *plength, tamcmc_alm_fn alm, void *alm_user, int capacity,
                             double *row_out, int *nmodes_out);

/* ---- red-giant models: mixed-mode host expander (BASELINE configs C1 / C4) ----
 * Replaces: the host half of model_RGB_asympt_aj_AppWidth_HarveyLike_v4 (model_id 25, models.cpp:4684-4927) and of
 * model_RGB_asympt_aj_CteWidth_HarveyLike_v4 (model_id 27, models.cpp:4334-4556): params / plength in those models' own layout ->
 * ONE mode-table row of `capacity` modes (row_out: TAMCMC_MT_HEADER + Nnoise + TAMCMC_MT_STRIDE * capacity doubles) for a context
 * created with plength = [capacity, 1 (step = x[2]-x[1]), 0,...,0, Nnoise, 0, 0].  step = x[2] - x[1] of the spectrum
 * (models.cpp:4714): the grid resolution of the ARMM solver.  *nmodes_out = number of modes (the number of l=1 mixed modes varies
 * from chain to chain); TAMCMC_ERR_ARG with *nmodes_out set when it exceeds `capacity`; TAMCMC_ERR_NONFINITE where the reference
 * exits (fmin - Dnu < 0, negative p-mode frequency) or finds no mixed mode. */
int tamcmc_host_expand_rgb_v4(int model_id, const double *params, const int *plength, double step, int capacity, double *row_out,
                              int *nmodes_out);
/* Replaces: solve_mm_asymptotic_O2from_l0 / solve_mm_asymptotic_O2p (external/ARMM/solver_mm.cpp:624-746, 470-604) with sigma_p = 0
 * and returns_pg_freqs = true: mixed-mode frequencies nu_m[*n_m] (sorted, duplicates within 2 resol removed), the pure p modes
 * nu_p[*n_p] with their local large separations dnup[*n_p], the pure g modes nu_g[*n_g].  Arrays of capacity `cap` (nu_p, dnup,
 * nu_g and the counts may be NULL). */
int tamcmc_host_armm_solve_from_l0(const double *nu_l0, int n_l0, int el, double delta0l, double DPl, double alpha, double q, double resol,
                                   double freq_min, double freq_max, int cap, double *nu_m, int *n_m, double *nu_p, double *dnup, int *n_p,
                                   double *nu_g, int *n_g);
int tamcmc_host_armm_solve_O2p(double Dnu_p, double epsilon, int el, double delta0l, double alpha_p, double nmax, double DPl, double alpha,
                               double q, double fmin, double fmax, double resol, int cap, double *nu_m, int *n_m, double *nu_p, double *dnup,
                               int *n_p, double *nu_g, int *n_g);
/* Replaces: tk::spline with second-derivative-zero boundaries (external/spline/src/spline.h), the bias of the l=1 mixed modes
 * (models.cpp:4834-4843): type 1 = cspline, 2 = cspline_hermite; out[i] = spline(xq[i]) (linear extrapolation outside). */
int tamcmc_host_spline_eval(const double *x, const double *y, int n, int type, const double *xq, int nq, double *out);

/* ---- red-giant models: the same expander with its two heavy loops on the device, all chains of a step in one call ----
 * Replaces: the per-chain OpenMP fan-out of generate_model (model_def.cpp:466-482, MALA.cpp:648) entering
 * model_RGB_asympt_aj_{AppWidth,CteWidth}_HarveyLike_v4 (models.cpp:4684-4927, 4334-4556) once per chain.  The (p mode, g mode) pair loop
 * of the mixed-mode solver (external/ARMM/solver_mm.cpp:558-573, 326-449) and the normalisation of the zeta function
 * (external/ARMM/bump_DP.cpp:126-163) -- 99 % of the host expander's time -- run as two kernels over the chains of the step; the rest is the
 * host code of tamcmc_host_expand_rgb_v4.  The pair loop reproduces the host solver's operations without FMA contraction, with tan / atan
 * evaluated in double-double and rounded once: frequencies equal the host expander's except where glibc's tan / atan are not correctly
 * rounded AND the solution sits on a rounding boundary (then 1 ulp); heights, widths and splittings agree to ~1e-15 (device cosines).
 * params: [nchains][params_stride] vectors in the layout of models 25 / 27; rows_out: [nchains][row_stride] mode-table rows of `capacity`
 * modes (normally the staging block of the evaluation context, tamcmc_gpu_params_staging); nmodes_out[c] (may be NULL); status_out[c] = what
 * tamcmc_host_expand_rgb_v4 returns for that chain; path_out[c] (may be NULL): 0 = device solve, otherwise the chain was handed to the host
 * solver of this library (flag bits of csrc/rgb_solver.cuh: an exact zero on a grid point, an unexpected shape, ...; -1: not exportable).
 * No CUDA device: TAMCMC_ERR_CUDA from tamcmc_gpu_rgb_create. */
typedef struct tamcmc_gpu_rgb tamcmc_gpu_rgb;
int tamcmc_gpu_rgb_create(tamcmc_gpu_rgb **out, int device, int max_chains);
void tamcmc_gpu_rgb_destroy(tamcmc_gpu_rgb *h);
int tamcmc_gpu_rgb_expand(tamcmc_gpu_rgb *h, int model_id, const double *params, int params_stride, const int *plength, double step,
                          int nchains, int capacity, double *rows_out, int row_stride, int *nmodes_out, int *status_out, int *path_out);
/* host-clock milliseconds of the last tamcmc_gpu_rgb_expand: prepare (host), device (copies, two kernels, sync), finish (host), total */
void tamcmc_gpu_rgb_timings(const tamcmc_gpu_rgb *h, double out[4]);
/* chain set-ups asked for since create, and how many of them were handed to the host solver (path_out != 0) */
void tamcmc_gpu_rgb_counts(const tamcmc_gpu_rgb *h, long *chains_total, long *chains_host);
const char *tamcmc_gpu_rgb_last_error(void);
/* TEST HOOK, no GPU: the device solver's segment decomposition (csrc/rgb_solver.cuh) run on the host for one chain, with glibc's tan / atan
 * (exact_trig = 0: reproduces tamcmc_host_expand_rgb_v4 bit for bit) or the double-double ones the device uses (1).  *flags_out: the flag
 * bits that would have sent the chain to the host solver.  Never called by the product path. */
int tamcmc_host_rgb_expand_emulated(int model_id, const double *params, const int *plength, double step, int capacity, double *row_out,
                                    int *nmodes_out, int exact_trig, int *flags_out);

/* ---- Alm from the precomputed grids (what the reference's ajAlm model actually uses) ----
 * Replaces: loadAllData + flatten_grid + init_2dgrid as run once by Config::Config (config.cpp:77-147;
 * external/Alm/Alm_cpp/Alm_interpol.cpp:11-79, bilinear_interpol.cpp:27-124): reads <grid_dir>/{gate,triangle}/A<l><m+l>.gz
 * (l = 1..3, m = 0..l; gzip'ed text "x=.." / "y=.." / "z=" / rows) and prepares the bicubic interpolators.  grid_dir is the
 * reference's external/Alm/data/Alm_grids_CPP/1deg_grids (or any directory written by tamcmc_alm_grids_make). */
typedef struct tamcmc_alm_grids tamcmc_alm_grids;
int tamcmc_alm_grids_load(const char *grid_dir, tamcmc_alm_grids **out);
void tamcmc_alm_grids_free(tamcmc_alm_grids *grids);
/* Replaces: Alm_interp_iter_preinitialised (Alm_interpol.cpp:188-348) -> interpolate_core -> gsl_interp2d_eval_e with
 * gsl_interp2d_bicubic (bilinear_interpol.cpp:115-135).  Has the tamcmc_alm_fn signature: pass it with `grids` as the user
 * pointer to tamcmc_host_expand_ajAlm.  theta0, delta in radians; outside the grid GSL raises GSL_EDOM (the reference's
 * process aborts): NaN here; l outside 1..3 or |m| > l: -9998 like the reference. */
double tamcmc_alm_grids_eval(int l, int m, double theta0, double delta, int filter_code, void *grids);
/* grid introspection (parity tests): axes and node values of grid (filter_code, l, |m|); z is [ny][nx] row-major */
int tamcmc_alm_grids_shape(const tamcmc_alm_grids *grids, int filter_code, int l, int m, int *nx, int *ny);
int tamcmc_alm_grids_nodes(const tamcmc_alm_grids *grids, int filter_code, int l, int m, double *x, double *y, double *z);
/* Replaces: the reference's GridMaker (external/Alm/Alm_cpp/do_grids.cpp:63-76, make_grids.cpp:17-131, gzip_compress.cpp:41-70):
 * writes <out_dir>/<gate|triangle>/A<l><m+l>.gz for l = 1..lmax, m = -l..l in the reference's text format (6 significant digits).
 * The shipped 1-degree grids are resol = pi/180, theta0 in [0, pi/2], delta in [0, pi/4]. */
int tamcmc_alm_grids_make(const char *out_dir, int filter_code, int lmax, double resol, double theta_min, double theta_max,
                          double delta_min, double delta_max);
const char *tamcmc_alm_grids_last_error(void);

const char *tamcmc_gpu_strerror(int status);
const char *tamcmc_gpu_last_error(void);
int tamcmc_gpu_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif
