// mcmc_driver.hpp -- the CALLER of the hot path: a fixed-seed restatement of the reference's adaptive
// Metropolis + parallel-tempering loop, restructured around ONE batched likelihood evaluation per step
// (INTEGRATION.md section 3).  Header-only C++17, no Eigen.
//
// Reference (tamcmc/sources/MALA.cpp):
//   constructor / Tcoefs = lambda^m                      :55-107
//   init_proposal (cov0 = diag(err^2), sigma, mu)        :246-292
//   new_prop_values (Cholesky of (cov+eps2)*sigma)       :339-369
//   update_position_MH (Metropolis ratio, NaN -> reject) :463-553
//   update_proposal (Robbins-Monro: mu, cov, sigma)      :296-319  with p1/p2/p3 projections :135-177
//   parallel_tempering (adjacent swap, tempered logL)    :397-461
//   execute (gamma = c0/(1+i), learning windows, mixing) :623-745
//
// Differences from the reference, all deliberate and documented in DESIGN.md:
//   * propose-all -> one evaluation of all chains -> accept-all (chains are independent within a step);
//   * one seeded std::mt19937_64 drives every draw on the calling thread (the reference seeds from time(NULL) and
//     races on the shared generator inside its OpenMP loop, MALA.cpp:62,648);
//   * priors are a user callback (priors_calc.cpp stays in the reference; host/priors.hpp restates the generic ones);
//     -inf prior -> the chain is masked out exactly like model_def.cpp:469-480;
//   * the normal draws of step i+1 and their products with the proposal's Cholesky factors are prepared WHILE the GPU evaluates
//     step i (they do not depend on its outcome; a factor that the learning update of step i changes is re-applied to the same
//     draws), so outside the learning windows the host half of a step hides behind the evaluation.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <algorithm>
#include <limits>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

namespace tamcmc {

struct DriverConfig {
    int Nchains = 10;
    double lambda_temp = 1.7;                 // Tcoefs[m] = lambda^m
    double c0 = 10.0, epsilon1 = 1e-12, epsi2 = 1e-10, A1 = 1e14, target_acceptance = 0.234;   // config_default.cfg:11-16 (the shipped
                                              // file has epsilon2 = 1e-12; formats.py:mala_config reads a run's own values)
    std::vector<long> Nt_learn = {1000, 1500, 100000};       // config_default.cfg:17
    std::vector<long> periods_learn = {1, 1};                 // config_default.cfg:18
    long dN_mixing = 1;                                       // config_default.cfg:28
    std::uint64_t seed = 1;
};

// Evaluator: int(const double* params [Nchains][stride], const unsigned char* active [Nchains], double* logL [Nchains])
//            returns 0, or a status; NaN logL is data (MALA.cpp:490,522).
using Evaluator = std::function<int(const double*, const unsigned char*, double*)>;
using Prior = std::function<double(const double* params_row)>;      // log prior of one full parameter vector
// the same evaluation in two halves (tamcmc_gpu_eval_begin / _end): the driver prepares the next proposal in between
struct AsyncEvaluator {
    std::function<int(const double*, const unsigned char*)> begin;
    std::function<int(double*)> end;
};

// Evaluator status codes that are DATA (some chain's logL is NaN and its proposal is rejected, MALA.cpp:490,522): TAMCMC_OK,
// TAMCMC_ERR_WINDOW, TAMCMC_ERR_NONFINITE (include/tamcmc_gpu.h).  Anything else (bad argument, unknown model, CUDA error)
// means the output buffer may not have been written at all.
inline bool eval_status_is_fatal(int rc) { return !(rc == 0 || rc == 4 || rc == 5); }

class Driver {
public:
    // state with the reference's names (model_def.h:54-66, MALA.h)
    std::vector<double> Tcoefs, sigma;                    // [Nchains]
    std::vector<double> mu, vars;                         // [Nchains][Nvars]
    std::vector<double> covarmat;                         // [Nchains][Nvars][Nvars]
    std::vector<double> params;                           // [Nchains][stride]
    std::vector<double> logLikelihood, logPrior, logPosterior, Pmove;
    std::vector<int> moved;
    long n_swap_tried = 0, n_swap_done = 0, n_eval_calls = 0;
    std::vector<long> n_accept;
    int last_eval_status = 0;                              // what the evaluator returned for the last step (0 = OK)

    // defer_initial_eval: the caller evaluates `params` itself (BatchDriver: one launch for all stars) and reports the result
    // through set_initial_logL()
    Driver(const DriverConfig& cfg_, int Nparams_, int stride_, const std::vector<double>& params0, const std::vector<int>& relax_index,
           const std::vector<double>& errors, Evaluator ev, Prior pr, bool defer_initial_eval = false)
        : cfg(cfg_), Nchains(cfg_.Nchains), Nparams(Nparams_), stride(stride_), Nvars((int)relax_index.size()), index_to_relax(relax_index),
          eval(std::move(ev)), prior(std::move(pr)), rng(cfg_.seed)
    {
        Tcoefs.resize(Nchains); sigma.resize(Nchains);
        for (int m = 0; m < Nchains; m++) Tcoefs[m] = std::pow(cfg.lambda_temp, m);                         // MALA.cpp:98-99
        mu.assign((size_t)Nchains * Nvars, 0.0); vars = mu;
        covarmat.assign((size_t)Nchains * Nvars * Nvars, 0.0);
        params.assign((size_t)Nchains * stride, 0.0);
        for (int m = 0; m < Nchains; m++) {
            for (int k = 0; k < Nparams; k++) params[(size_t)m * stride + k] = params0[(size_t)k];
            for (int v = 0; v < Nvars; v++) {
                vars[(size_t)m * Nvars + v] = params0[(size_t)index_to_relax[v]];
                mu[(size_t)m * Nvars + v] = vars[(size_t)m * Nvars + v];                                      // MALA.cpp:286
                covarmat[((size_t)m * Nvars + v) * Nvars + v] = errors[v] * errors[v];                         // MALA.cpp:270-276
            }
            sigma[m] = std::pow(2.38, 2) * std::pow(Tcoefs[m], 0.2) / Nvars;                                   // MALA.cpp:277
        }
        logLikelihood.assign(Nchains, 0.0); logPrior = logLikelihood; logPosterior = logLikelihood; Pmove = logLikelihood;
        moved.assign(Nchains, 0); n_accept.assign(Nchains, 0);
        prop_params = params; prop_vars = vars; prop_logL = logLikelihood; prop_logPrior = logLikelihood; active.assign(Nchains, 1);
        // initial model (Model_def constructor, model_def.cpp:142-147)
        for (int m = 0; m < Nchains; m++) { logPrior[m] = prior(&params[(size_t)m * stride]); active[m] = std::isinf(logPrior[m]) ? 0 : 1; }
        if (!defer_initial_eval) {
            evaluate_current("initial model");
            set_initial_logL(logLikelihood.data());
        }
    }

    void set_initial_logL(const double* logL)
    {
        for (int m = 0; m < Nchains; m++) {
            logLikelihood[m] = logL[m];
            logPosterior[m] = active[m] ? logLikelihood[m] + logPrior[m] : -std::numeric_limits<double>::infinity();
        }
    }

    // Restart from a previous run (the reference's do_restore_proposal / do_restore_variables switches; the values come from
    // its three restore files, tamcmc-c_b200/formats.py:read_restore).
    // MALA::restore_proposal (MALA.cpp:191-246): scale, mean and covariance of every chain's proposal law.
    void restore_proposal(const double* sigma_in /*[Nchains]*/, const double* mu_in /*[Nchains][Nvars]*/,
                          const double* covarmat_in /*[Nchains][Nvars][Nvars]*/)
    {
        std::copy(sigma_in, sigma_in + Nchains, sigma.begin());
        std::copy(mu_in, mu_in + (size_t)Nchains * Nvars, mu.begin());
        std::copy(covarmat_in, covarmat_in + (size_t)Nchains * Nvars * Nvars, covarmat.begin());
        if (!chol_dirty.empty()) chol_dirty.assign((size_t)Nchains, 1);       // every factor is stale; prepared draws get it re-applied
    }
    // Chain positions of a previous run (Model_def constructed on the restored vars, model_def.cpp:142-147): priors and
    // likelihoods are evaluated again at the restored position.
    void restore_variables(const double* vars_in /*[Nchains][Nvars]*/)
    {
        std::copy(vars_in, vars_in + (size_t)Nchains * Nvars, vars.begin());
        for (int m = 0; m < Nchains; m++) {
            for (int v = 0; v < Nvars; v++) params[(size_t)m * stride + index_to_relax[v]] = vars[(size_t)m * Nvars + v];
            logPrior[m] = prior(&params[(size_t)m * stride]);
            active[m] = std::isinf(logPrior[m]) ? 0 : 1;
        }
        prop_params = params; prop_vars = vars;
        evaluate_current("restored position");
        set_initial_logL(logLikelihood.data());
    }

    // buffers of the two-phase interface (propose -> caller evaluates -> finish)
    const double* proposal_params() const { return prop_params.data(); }
    const unsigned char* active_mask() const { return active.data(); }
    double* proposal_logL() { return prop_logL.data(); }
    int row_stride() const { return stride; }

    // one iteration i of MALA::execute (MALA.cpp:646-700).  Returns the evaluator's status.  The proposal's log-likelihoods
    // are preset to NaN, so a chain the evaluator did not write (a failed call) is rejected like any NaN likelihood
    // (MALA.cpp:522-524) instead of being judged on the previous step's values; a fatal status (eval_status_is_fatal) is
    // returned to the caller, who decides whether to go on.
    int step(long i)
    {
        propose(i);
        // ---- ONE batched evaluation (was: generate_model per chain inside the OpenMP loop); the draws of the next step are
        // prepared while it runs (same order of random numbers with and without an asynchronous evaluator) ----
        std::fill(prop_logL.begin(), prop_logL.end(), std::numeric_limits<double>::quiet_NaN());
        int rc;
        if (async.begin) {
            rc = async.begin(prop_params.data(), active.data());
            prepare_next();
            if (!eval_status_is_fatal(rc)) rc = async.end(prop_logL.data());
        } else {
            rc = eval(prop_params.data(), active.data(), prop_logL.data());
            prepare_next();
        }
        n_eval_calls++;
        last_eval_status = rc;
        if (eval_status_is_fatal(rc)) std::fill(prop_logL.begin(), prop_logL.end(), std::numeric_limits<double>::quiet_NaN());
        finish(i);
        return rc;
    }

    void set_async_evaluator(AsyncEvaluator a) { async = std::move(a); }

    // Between the two halves of an iteration: the N(0, I) draws of the NEXT proposal (serially, in chain order) and L z for
    // every chain whose factor is current.  propose() re-applies a factor that finish() changes in between.
    void prepare_next()
    {
        const int n = Nvars;
        if (chol_all.empty()) { chol_all.assign((size_t)Nchains * n * n, 0.0); chol_dirty.assign((size_t)Nchains, 1); }
        z_next.resize((size_t)Nchains * n);
        lz_next.resize((size_t)Nchains * n);
        lz_valid.assign((size_t)Nchains, 0);
        for (size_t k = 0; k < z_next.size(); k++) z_next[k] = normal01();
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1) if (Nvars >= 32)
#endif
        for (int m = 0; m < Nchains; m++)
            if (!chol_dirty[(size_t)m]) { apply_factor(m, &z_next[(size_t)m * n], &lz_next[(size_t)m * n]); lz_valid[(size_t)m] = 1; }
        have_next = true;
    }

    // first half of an iteration: new positions and their priors for every chain
    void propose(long i)
    {
        gamma = cfg.c0 / (1.0 + (double)i);                                                                   // MALA.cpp:646
        // ---- propose all chains (MALA.cpp:481-486).  Random numbers are drawn serially in chain order (deterministic
        // whatever the thread count); factorisations, matrix-vector products and priors then run one chain per thread ----
        const int n = Nvars;
        if (!have_next) prepare_next();                        // first iteration (or a caller that drives propose()/finish() itself)
        z_all.swap(z_next); lz_all.swap(lz_next); lz_ok.swap(lz_valid);
        have_next = false;
        nonfinite.assign((size_t)Nchains, 0);
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1) if (Nvars >= 32)
#endif
        for (int m = 0; m < Nchains; m++) {
            // a factor that changed since the draws were prepared (learning update of the previous step) is applied now
            if (!lz_ok[(size_t)m] || chol_dirty[(size_t)m]) apply_factor(m, &z_all[(size_t)m * n], &lz_all[(size_t)m * n]);
            nonfinite[(size_t)m] = move_from(m, &lz_all[(size_t)m * n]) ? 0 : 1;
        }
        for (int m = 0; m < Nchains; m++)                       // MALA.cpp:356-366: redraw until finite (never seen in practice)
            for (int tries = 0; nonfinite[(size_t)m] && tries < 8; tries++) {
                for (int a = 0; a < n; a++) z_all[(size_t)m * n + a] = normal01();
                apply_factor(m, &z_all[(size_t)m * n], &lz_all[(size_t)m * n]);
                nonfinite[(size_t)m] = move_from(m, &lz_all[(size_t)m * n]) ? 0 : 1;
            }
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1) if (Nvars >= 32)
#endif
        for (int m = 0; m < Nchains; m++) {
            for (int k = 0; k < Nparams; k++) prop_params[(size_t)m * stride + k] = params[(size_t)m * stride + k];
            for (int v = 0; v < Nvars; v++) prop_params[(size_t)m * stride + index_to_relax[v]] = prop_vars[(size_t)m * Nvars + v];
            prop_logPrior[m] = prior(&prop_params[(size_t)m * stride]);
            active[m] = (prop_logPrior[m] == -std::numeric_limits<double>::infinity()) ? 0 : 1;               // model_def.cpp:469
        }
    }

    // second half: proposal_logL() holds the tempered log-likelihoods of proposal_params()
    void finish(long i)
    {
        // ---- accept / reject (MALA.cpp:490-548), learn (MALA.cpp:656-668) ----
        u_all.resize((size_t)Nchains);
        for (int m = 0; m < Nchains; m++) u_all[(size_t)m] = uniform01();
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1) if (Nvars >= 32)
#endif
        for (int m = 0; m < Nchains; m++) {
            const double u = u_all[(size_t)m];
            double r;
            const double prop_post = active[m] ? prop_logL[m] + prop_logPrior[m] : -std::numeric_limits<double>::infinity();
            if (!active[m]) r = 0.0;
            else if (std::isnan(prop_logL[m])) r = 0.0;
            else { r = std::exp(prop_post - logPosterior[m]); if (r > 1.0) r = 1.0; if (std::isnan(r)) r = 0.0; }
            if (u <= r && active[m] && !std::isnan(prop_logL[m])) {
                for (int k = 0; k < Nparams; k++) params[(size_t)m * stride + k] = prop_params[(size_t)m * stride + k];
                for (int v = 0; v < Nvars; v++) vars[(size_t)m * Nvars + v] = prop_vars[(size_t)m * Nvars + v];
                logLikelihood[m] = prop_logL[m]; logPrior[m] = prop_logPrior[m]; logPosterior[m] = prop_post;
                moved[m] = 1; n_accept[m]++;
            } else moved[m] = 0;
            Pmove[m] = r;
            int learn = 0, which = 0;
            for (size_t l = 0; l + 1 < cfg.Nt_learn.size() && l < cfg.periods_learn.size(); l++)
                if (i >= cfg.Nt_learn[l] && i < cfg.Nt_learn[l + 1]) { learn = 1; which = (int)l; }
            if (learn && (i % cfg.periods_learn[(size_t)which]) == 0) update_proposal(m, Pmove[m]);
        }
        // ---- parallel tempering (MALA.cpp:688-700, 397-461) ----
        if (Nchains > 1 && cfg.dN_mixing > 0 && i % cfg.dN_mixing == 0 && i != 0) parallel_tempering();
    }

    int n_vars() const { return Nvars; }
    int n_chains() const { return Nchains; }

private:
    // likelihoods of the CURRENT positions (constructor, restore): there is nothing to fall back on if this fails
    void evaluate_current(const char* what)
    {
        std::fill(logLikelihood.begin(), logLikelihood.end(), std::numeric_limits<double>::quiet_NaN());
        const int rc = eval(params.data(), active.data(), logLikelihood.data());
        n_eval_calls++;
        last_eval_status = rc;
        if (eval_status_is_fatal(rc)) throw std::runtime_error(std::string("tamcmc::Driver: evaluation of the ") + what + " failed with status " + std::to_string(rc));
    }

    DriverConfig cfg;
    int Nchains, Nparams, stride, Nvars;
    std::vector<int> index_to_relax;
    Evaluator eval;
    Prior prior;
    std::mt19937_64 rng;
    double gamma = 0.0;
    std::vector<double> prop_params, prop_vars, prop_logL, prop_logPrior, chol_all, z_all, lz_all, z_next, lz_next;
    std::vector<unsigned char> lz_ok, lz_valid;
    bool have_next = false;
    AsyncEvaluator async;
    std::vector<unsigned char> chol_dirty, nonfinite;
    std::vector<double> u_all;
    std::vector<unsigned char> active;

    double uniform01() { return std::generate_canonical<double, 53>(rng); }
    double normal01()
    {   // Box-Muller on the seeded generator: the same sequence on every platform
        double u1 = uniform01();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925286766559 * uniform01());
    }

    // MALA.cpp:339-369: ran = vars + chol((covarmat + epsilon2) * sigma) * N(0, I).  The reference factorises at every
    // call; chol((C + eps2) sigma) = sqrt(sigma) chol(C + eps2), so the factor of C + eps2 is cached per chain and only
    // recomputed after update_proposal changed C (with the GPU likelihood the factorisation would otherwise dominate a step).
    // apply_factor: out = L z (refreshing L first if C changed); move_from: prop_vars = vars + sqrt(sigma) (L z).
    void apply_factor(int m, const double* zm, double* out)
    {
        const int n = Nvars;
        double* L = &chol_all[(size_t)m * n * n];
        if (chol_dirty[(size_t)m]) {
            const double* C = &covarmat[(size_t)m * n * n];
            for (int a = 0; a < n; a++)
                for (int b = 0; b <= a; b++) {
                    const double* La = &L[(size_t)a * n];
                    const double* Lb = &L[(size_t)b * n];
                    double dot = 0.0;                       // rows of L are contiguous: a vectorised dot product (fixed lane
#ifdef _OPENMP                                              // order for a given build, so runs stay reproducible)
#pragma omp simd reduction(+ : dot)
#endif
                    for (int k = 0; k < b; k++) dot += La[k] * Lb[k];
                    const double s = C[(size_t)a * n + b] + (a == b ? cfg.epsi2 : 0.0) - dot;
                    L[(size_t)a * n + b] = (a == b) ? std::sqrt(s > 0 ? s : 0.0) : (L[(size_t)b * n + b] > 0 ? s / L[(size_t)b * n + b] : 0.0);
                }
            chol_dirty[(size_t)m] = 0;
        }
        for (int a = 0; a < n; a++) {
            const double* La = &L[(size_t)a * n];
            double s = 0.0;
#ifdef _OPENMP
#pragma omp simd reduction(+ : s)
#endif
            for (int k = 0; k <= a; k++) s += La[k] * zm[k];
            out[a] = s;
        }
    }
    bool move_from(int m, const double* lz)
    {
        const int n = Nvars;
        const double ssig = std::sqrt(sigma[m]);
        bool finite = true;
        for (int a = 0; a < n; a++) {
            const double s = vars[(size_t)m * n + a] + ssig * lz[a];
            prop_vars[(size_t)m * n + a] = s;
            finite = finite && std::isfinite(s);
        }
        return finite;
    }

    // MALA.cpp:296-319 with the projections p1/p2/p3 (MALA.cpp:135-177)
    void update_proposal(int m, double acceptance)
    {
        const int n = Nvars;
        double* mu_m = &mu[(size_t)m * n];
        double* C = &covarmat[(size_t)m * n * n];
        const double* v = &vars[(size_t)m * n];
        double nrm = 0.0;
        for (int a = 0; a < n; a++) { mu_m[a] = mu_m[a] + gamma * (v[a] - mu_m[a]); nrm += mu_m[a] * mu_m[a]; }
        nrm = std::sqrt(nrm);
        if (nrm > cfg.A1) for (int a = 0; a < n; a++) mu_m[a] *= cfg.A1 / nrm;
        double fro = 0.0;
        for (int a = 0; a < n; a++)
            for (int b = 0; b < n; b++) {
                const double mat = (v[a] - mu_m[a]) * (v[b] - mu_m[b]);
                C[(size_t)a * n + b] = C[(size_t)a * n + b] + gamma * (mat - C[(size_t)a * n + b]);
                fro += C[(size_t)a * n + b] * C[(size_t)a * n + b];
            }
        fro = std::sqrt(fro);
        if (fro > cfg.A1) for (size_t k = 0; k < (size_t)n * n; k++) C[k] *= cfg.A1 / fro;
        chol_dirty[(size_t)m] = 1;
        double s = sigma[m] + gamma * (acceptance - cfg.target_acceptance);
        if (s < cfg.epsilon1) s = cfg.epsilon1;
        if (s > cfg.A1) s = cfg.A1;
        sigma[m] = s;
    }

    // MALA.cpp:397-461: swap two adjacent chains; the stored likelihoods are TEMPERED.  (The reference computes logPosterior[B]
    // with logPrior[A] AFTER it has overwritten logPrior[A] by B's value, MALA.cpp:447; the intended prior is used here --
    // identical under the flat priors of the tests.)
    void parallel_tempering()
    {
        const double u = uniform01();
        const int A = (int)(rng() % (std::uint64_t)(Nchains - 1)), B = A + 1;      // random_int_vals(0, Nchains-1), MALA.cpp:179-189
        const double LA_TB = logLikelihood[A] * Tcoefs[A] / Tcoefs[B];
        const double LB_TA = logLikelihood[B] * Tcoefs[B] / Tcoefs[A];
        double r = std::exp(LA_TB + LB_TA - logLikelihood[A] - logLikelihood[B]);
        if (r > 1.0) r = 1.0;
        n_swap_tried++;
        if (u <= r) {
            for (int k = 0; k < stride; k++) std::swap(params[(size_t)A * stride + k], params[(size_t)B * stride + k]);
            for (int v = 0; v < Nvars; v++) std::swap(vars[(size_t)A * Nvars + v], vars[(size_t)B * Nvars + v]);
            const double prA = logPrior[A], prB = logPrior[B];
            logLikelihood[A] = LB_TA; logPrior[A] = prB; logPosterior[A] = LB_TA + prB;
            logLikelihood[B] = LA_TB; logPrior[B] = prA; logPosterior[B] = LA_TB + prA;
            std::swap(moved[A], moved[B]); std::swap(Pmove[A], Pmove[B]);
            n_swap_done++;
        }
    }
};

// Many independent stars (BASELINE config C5; the reference runs one process per star, scripts/slurm/job.sh): one Driver per
// star, all proposals of all stars evaluated by ONE batched call per step.  Stars are independent, so their host halves run one
// star per OpenMP thread; every Driver owns its seeded generator, so results do not depend on the thread count.
using BatchEvaluator = std::function<int(const double* /*[nstars][Nchains][stride]*/, const unsigned char* /*[nstars][Nchains]*/,
                                         double* /*[nstars][Nchains]*/)>;
class BatchDriver {
public:
    std::vector<std::unique_ptr<Driver>> stars;
    BatchDriver(std::vector<std::unique_ptr<Driver>> d, BatchEvaluator ev) : stars(std::move(d)), eval(std::move(ev))
    {
        const size_t S = stars.size();
        Nchains = stars[0]->n_chains(); stride = stars[0]->row_stride();
        P.assign(S * Nchains * stride, 0.0); act.assign(S * Nchains, 1); L.assign(S * Nchains, 0.0);
        // initial models of all stars in one launch (Model_def constructor, model_def.cpp:142-147)
        for (size_t s = 0; s < S; s++) {
            std::copy(stars[s]->params.begin(), stars[s]->params.end(), P.begin() + s * Nchains * stride);
            std::copy(stars[s]->active_mask(), stars[s]->active_mask() + Nchains, act.begin() + s * Nchains);
        }
        std::fill(L.begin(), L.end(), std::numeric_limits<double>::quiet_NaN());
        last_eval_status = eval(P.data(), act.data(), L.data());
        if (eval_status_is_fatal(last_eval_status))
            throw std::runtime_error("tamcmc::BatchDriver: evaluation of the initial models failed with status " + std::to_string(last_eval_status));
        for (size_t s = 0; s < S; s++) stars[s]->set_initial_logL(&L[s * Nchains]);
    }
    int last_eval_status = 0;
    // returns the evaluator's status; unwritten / failed evaluations reject every proposal (NaN), see Driver::step
    int step(long i)
    {
        const long S = (long)stars.size();
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
        for (long s = 0; s < S; s++) {
            stars[(size_t)s]->propose(i);
            std::copy(stars[(size_t)s]->proposal_params(), stars[(size_t)s]->proposal_params() + (size_t)Nchains * stride, P.begin() + (size_t)s * Nchains * stride);
            std::copy(stars[(size_t)s]->active_mask(), stars[(size_t)s]->active_mask() + Nchains, act.begin() + (size_t)s * Nchains);
        }
        std::fill(L.begin(), L.end(), std::numeric_limits<double>::quiet_NaN());
        int rc = async.begin ? async.begin(P.data(), act.data()) : eval(P.data(), act.data(), L.data());
        // the next step's draws of every star, while the batched evaluation runs (same order with a synchronous evaluator)
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
        for (long s = 0; s < S; s++) stars[(size_t)s]->prepare_next();
        if (async.begin && !eval_status_is_fatal(rc)) rc = async.end(L.data());
        last_eval_status = rc;
        if (eval_status_is_fatal(rc)) std::fill(L.begin(), L.end(), std::numeric_limits<double>::quiet_NaN());
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
        for (long s = 0; s < S; s++) {
            std::copy(L.begin() + (size_t)s * Nchains, L.begin() + (size_t)(s + 1) * Nchains, stars[(size_t)s]->proposal_logL());
            stars[(size_t)s]->finish(i);
        }
        return rc;
    }
    void set_async_evaluator(AsyncEvaluator a) { async = std::move(a); }
private:
    BatchEvaluator eval;
    AsyncEvaluator async;
    int Nchains = 0, stride = 0;
    std::vector<double> P, L;
    std::vector<unsigned char> act;
};

}  // namespace tamcmc
