// model_def_gpu.hpp -- host-side C++ mirror of the hot-path half of the reference's `Model_def`
// (tamcmc/headers/model_def.h:22-93, tamcmc/sources/model_def.cpp) on top of the C ABI of
// include/tamcmc_gpu.h.  Header-only, no Eigen, no torch: plain std::vector storage with the reference's
// member names and meanings, so a maintainer can swap it in and the C++ tests read like
// the reference's own call sites.
//
//   reference                                             | here
//   ------------------------------------------------------+-----------------------------------------------
//   Model_def(Config*, Tcoefs, verbose)  model_def.cpp:28  | ModelDefGPU(model_fct_name_switch, plength, x, y, Nchains, Tcoefs, p)
//   generate_model(data, m, Tcoefs)      model_def.cpp:466 | generate_models(logPrior[]) : ALL chains m in one launch
//   call_model_explicit(...)             model_def.cpp:209 | call_model_explicit(params) -> model spectrum
//   logLikelihood[m] (tempered)          model_def.h:59    | logLikelihood[m]
//   exit(EXIT_FAILURE)                   several           | tamcmc_error (status code + text); NaN logL is data
//
// There is no CPU path here: every method forwards to libtamcmc_gpu.so and throws if it fails.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/tamcmc_gpu.h"

namespace tamcmc {

struct tamcmc_error : std::runtime_error {
    int status;
    tamcmc_error(int s, const std::string& what) : std::runtime_error(what), status(s) {}
};

inline void check(int rc, const char* where)
{
    if (rc == TAMCMC_OK) return;
    std::string msg = std::string(where) + ": " + tamcmc_gpu_strerror(rc);
    if (rc == TAMCMC_ERR_CUDA) msg += std::string(" -- ") + tamcmc_gpu_last_error();
    throw tamcmc_error(rc, msg);
}

// One star/slice = the reference's Data{x, y, Nx} (tamcmc/headers/data.h) + the model selection of its Model_def.
struct StarData {
    int model_fct_name_switch;          // Config/default/models_ctrl.list id (model_def.cpp:220-388)
    std::vector<int> plength;           // 11 entries (io_ms_global.cpp:1315-1325)
    int Nparams;
    std::vector<double> x, y;
    std::vector<double> sigma_y;        // Data.sigma_y, only read by the chi_square likelihood; empty = ones (config.cpp:367-374)
};

class ModelDefGPU {
public:
    // public state with the reference's names (model_def.h:54-66)
    std::vector<double> params;          // [nstars][Nmodels][Nparams_stride] row-major (reference: MatrixXd params(Nmodels, Nparams))
    std::vector<double> logLikelihood;   // [nstars][Nmodels], tempered (model_def.cpp:401)
    std::vector<double> init_logLikelihood;
    std::vector<double> logPrior;        // filled by the caller's call_prior (stays on the host, out of scope here)
    std::vector<double> logPosterior;
    std::vector<int> status;             // TAMCMC_CHAIN_* bits per chain

    ModelDefGPU(const std::vector<StarData>& stars, int Nmodels_, const std::vector<double>& Tcoefs, double likelihood_params = 1.0,
                int likelihood_fct_name_switch = TAMCMC_LIKELIHOOD_CHI22P, int device = 0)
        : Nmodels(Nmodels_), nstars((int)stars.size())
    {
        if ((int)Tcoefs.size() != Nmodels) throw tamcmc_error(TAMCMC_ERR_ARG, "Tcoefs.size() != Nmodels");
        std::vector<tamcmc_gpu_star> s(stars.size());
        for (size_t i = 0; i < stars.size(); i++) {
            if (stars[i].plength.size() != 11 || stars[i].x.size() != stars[i].y.size())
                throw tamcmc_error(TAMCMC_ERR_ARG, "plength must have 11 entries and x, y the same length");
            s[i] = tamcmc_gpu_star();
            s[i].model_id = stars[i].model_fct_name_switch;
            for (int k = 0; k < 11; k++) s[i].plength[k] = stars[i].plength[k];
            s[i].Nparams = stars[i].Nparams;
            s[i].x = stars[i].x.data(); s[i].y = stars[i].y.data(); s[i].N = (long)stars[i].x.size();
            s[i].sigma_y = (stars[i].sigma_y.size() == stars[i].x.size()) ? stars[i].sigma_y.data() : nullptr;
            Nx.push_back((long)stars[i].x.size());
            Nparams_of.push_back(stars[i].Nparams);
        }
        check(tamcmc_gpu_create(device, nstars, s.data(), Nmodels, Tcoefs.data(), likelihood_params, likelihood_fct_name_switch, &ctx),
              "tamcmc_gpu_create");
        Nparams_stride = tamcmc_gpu_params_stride(ctx);
        const size_t n = (size_t)nstars * Nmodels;
        params.assign(n * Nparams_stride, 0.0);
        logLikelihood.assign(n, std::numeric_limits<double>::quiet_NaN());
        init_logLikelihood = logLikelihood;
        logPrior.assign(n, 0.0);
        logPosterior.assign(n, -std::numeric_limits<double>::infinity());
        status.assign(n, 0);
        active.assign(n, 1);
    }
    ModelDefGPU(const ModelDefGPU&) = delete;
    ModelDefGPU& operator=(const ModelDefGPU&) = delete;
    ~ModelDefGPU() { tamcmc_gpu_destroy(ctx); }

    double* params_row(int star, int m) { return params.data() + ((size_t)star * Nmodels + m) * Nparams_stride; }
    int params_stride() const { return Nparams_stride; }
    int n_models() const { return Nmodels; }
    int n_stars() const { return nstars; }

    // generate_model for ALL chains at once (model_def.cpp:466-482): chains whose logPrior is -inf are not evaluated and
    // take init_logLikelihood / logPosterior = -inf exactly like the reference; the others get the tempered logL.
    // Returns the C-ABI status (TAMCMC_OK, or ERR_WINDOW / ERR_NONFINITE when some chain was flagged: NaN logL is data,
    // MALA.cpp:490,522).  Throws on CUDA / argument errors.
    // What the reference's constructor does with the initial parameters (model_def.cpp:142-153): model + likelihood of EVERY
    // chain "whatever the situation" (no prior short-circuit), kept as init_logLikelihood -- the value generate_model hands
    // back for a chain whose prior is -inf (model_def.cpp:476-480).  Call it once `params` holds the initial vectors;
    // generate_models() / generate_models_begin() run it on their first call otherwise.
    int initialise()
    {
        const size_t n = (size_t)nstars * Nmodels;
        std::fill(active.begin(), active.end(), (unsigned char)1);
        const int rc = tamcmc_gpu_eval(ctx, params.data(), active.data(), logLikelihood.data(), status.data());
        if (rc != TAMCMC_OK && rc != TAMCMC_ERR_WINDOW && rc != TAMCMC_ERR_NONFINITE) check(rc, "tamcmc_gpu_eval (initial models)");
        init_logLikelihood = logLikelihood;
        for (size_t i = 0; i < n; i++) logPosterior[i] = logLikelihood[i] + logPrior[i];
        initialised = true;
        return rc;
    }

    int generate_models()
    {
        if (!initialised) initialise();
        const size_t n = (size_t)nstars * Nmodels;
        for (size_t i = 0; i < n; i++) active[i] = (logPrior[i] != -std::numeric_limits<double>::infinity()) ? 1 : 0;
        const int rc = tamcmc_gpu_eval(ctx, params.data(), active.data(), logLikelihood.data(), status.data());
        if (rc != TAMCMC_OK && rc != TAMCMC_ERR_WINDOW && rc != TAMCMC_ERR_NONFINITE) check(rc, "tamcmc_gpu_eval");
        for (size_t i = 0; i < n; i++) {
            if (active[i]) logPosterior[i] = logLikelihood[i] + logPrior[i];
            else { logLikelihood[i] = init_logLikelihood[i]; logPosterior[i] = -std::numeric_limits<double>::infinity(); }
        }
        return rc;
    }

    // generate_models() in two halves (tamcmc_gpu_eval_begin / _end): the caller does host work that does not depend on the
    // likelihoods in between (e.g. the random numbers of the next proposal).  Same results as generate_models().
    void generate_models_begin()
    {
        if (!initialised) initialise();
        const size_t n = (size_t)nstars * Nmodels;
        for (size_t i = 0; i < n; i++) active[i] = (logPrior[i] != -std::numeric_limits<double>::infinity()) ? 1 : 0;
        check(tamcmc_gpu_eval_begin(ctx, params.data(), active.data()), "tamcmc_gpu_eval_begin");
    }
    int generate_models_end()
    {
        const size_t n = (size_t)nstars * Nmodels;
        const int rc = tamcmc_gpu_eval_end(ctx, logLikelihood.data(), status.data());
        if (rc != TAMCMC_OK && rc != TAMCMC_ERR_WINDOW && rc != TAMCMC_ERR_NONFINITE) check(rc, "tamcmc_gpu_eval_end");
        for (size_t i = 0; i < n; i++) {
            if (active[i]) logPosterior[i] = logLikelihood[i] + logPrior[i];
            else { logLikelihood[i] = init_logLikelihood[i]; logPosterior[i] = -std::numeric_limits<double>::infinity(); }
        }
        return rc;
    }

    // call_model_explicit (model_def.cpp:209-218): the model spectrum of one parameter vector
    std::vector<double> call_model_explicit(const std::vector<double>& params0, int star = 0)
    {
        if ((int)params0.size() < Nparams_of.at((size_t)star)) throw tamcmc_error(TAMCMC_ERR_ARG, "params0 shorter than Nparams");
        std::vector<double> m((size_t)Nx.at((size_t)star));
        check(tamcmc_gpu_model(ctx, star, params0.data(), m.data()), "tamcmc_gpu_model");
        return m;
    }

    tamcmc_gpu_ctx* handle() { return ctx; }

private:
    tamcmc_gpu_ctx* ctx = nullptr;
    int Nmodels, nstars, Nparams_stride = 0;
    std::vector<long> Nx;
    std::vector<int> Nparams_of;
    std::vector<unsigned char> active;
    bool initialised = false;          // initialise() has filled init_logLikelihood
};

}  // namespace tamcmc
