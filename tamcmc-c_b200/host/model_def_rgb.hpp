// model_def_rgb.hpp -- the red-giant models behind the reference's Model_def interface, with the set-up on the device.
//
// Mirrors, for model_RGB_asympt_aj_AppWidth_HarveyLike_v4 / ..._CteWidth_... (models_ctrl.list ids 25 / 27, tamcmc/sources/models.cpp:
// 4684-5079, 4334-4682), the hot-path half of Model_def (tamcmc/headers/model_def.h:22-93, tamcmc/sources/model_def.cpp:220-482):
//
//   reference                                              | here
//   -------------------------------------------------------+----------------------------------------------------------------------
//   Model_def(Config*, Tcoefs, verbose)   model_def.cpp:28  | ModelDefRGB(model_id, plength, x, y, Nmodels, Tcoefs, capacity, p)
//   params = MatrixXd(Nmodels, Nparams)   model_def.h:54    | params [Nmodels][Nparams] row-major, in the MODEL's own layout
//   generate_model(data, m, Tcoefs) x Nmodels   :466-482    | generate_models(): ONE tamcmc_gpu_rgb_expand (mixed-mode pair loop + zeta
//                                                           |   normalisation of all chains on the device, rows straight into the staging
//                                                           |   block) + ONE tamcmc_gpu_eval
//   call_model_explicit(params)           model_def.cpp:209 | call_model_explicit(params0) -> model spectrum
//   exit() in the model function (models.cpp:4852-4858 ...) | that chain's logLikelihood = NaN (a rejected proposal), expand_status[m] says why
//
// Header-only on top of include/tamcmc_gpu.h; no CPU path: both handles need a CUDA device (TAMCMC_ERR_CUDA otherwise).
#pragma once
#include "model_def_gpu.hpp"

namespace tamcmc {

class ModelDefRGB {
public:
    std::vector<double> params;            // [Nmodels][Nparams]: the reference's parameter vectors (ids 25 / 27 layout)
    std::vector<double> logLikelihood;     // [Nmodels], tempered (model_def.cpp:401)
    std::vector<double> init_logLikelihood;
    std::vector<double> logPrior, logPosterior;
    std::vector<int> status;               // TAMCMC_CHAIN_* bits of the evaluation
    std::vector<int> expand_status;        // per chain: what the model function's set-up returned (TAMCMC_OK, TAMCMC_ERR_NONFINITE where the reference exits)
    std::vector<int> expand_path;          // per chain: 0 = solved on the device, otherwise handed to the host solver of the library
    std::vector<int> nmodes;               // per chain: modes of the last set-up (the number of l=1 mixed modes varies)

    ModelDefRGB(int model_fct_name_switch, const std::vector<int>& plength_, const std::vector<double>& x, const std::vector<double>& y,
                int Nmodels_, const std::vector<double>& Tcoefs, int capacity_ = 160, double likelihood_params = 1.0, int device = 0)
        : model_id(model_fct_name_switch), plength(plength_), Nmodels(Nmodels_), capacity(capacity_), Nx((long)x.size())
    {
        if (plength.size() != 11 || x.size() != y.size() || x.size() < 3 || (int)Tcoefs.size() != Nmodels)
            throw tamcmc_error(TAMCMC_ERR_ARG, "plength must have 11 entries, x and y the same length, Tcoefs.size() == Nmodels");
        if (model_id != 25 && model_id != 27) throw tamcmc_error(TAMCMC_ERR_MODEL, "ModelDefRGB: model ids 25 / 27");
        Nparams = 0;
        for (int k = 0; k < 11; k++) Nparams += plength[(size_t)k];
        step = x[2] - x[1];                                            // models.cpp:4714
        const int Nnoise = plength[8];
        tamcmc_gpu_star s = tamcmc_gpu_star();
        s.model_id = TAMCMC_MODEL_MODE_TABLE;
        s.plength[0] = capacity; s.plength[1] = 1;                     // step = x[2] - x[1] like the red-giant models
        s.plength[8] = Nnoise;
        s.Nparams = TAMCMC_MT_HEADER + Nnoise + TAMCMC_MT_STRIDE * capacity;
        s.x = x.data(); s.y = y.data(); s.N = Nx;
        check(tamcmc_gpu_create(device, 1, &s, Nmodels, Tcoefs.data(), likelihood_params, TAMCMC_LIKELIHOOD_CHI22P, &ctx), "tamcmc_gpu_create");
        const int rc = tamcmc_gpu_rgb_create(&rgb, device, Nmodels);
        if (rc != TAMCMC_OK) { tamcmc_gpu_destroy(ctx); ctx = nullptr; throw tamcmc_error(rc, std::string("tamcmc_gpu_rgb_create: ") + tamcmc_gpu_rgb_last_error()); }
        rows = tamcmc_gpu_params_staging(ctx, &row_stride);
        const size_t n = (size_t)Nmodels;
        params.assign(n * (size_t)Nparams, 0.0);
        logLikelihood.assign(n, std::numeric_limits<double>::quiet_NaN());
        init_logLikelihood = logLikelihood;
        logPrior.assign(n, 0.0);
        logPosterior.assign(n, -std::numeric_limits<double>::infinity());
        status.assign(n, 0); expand_status.assign(n, 0); expand_path.assign(n, 0); nmodes.assign(n, 0);
        active.assign(n, 1);
    }
    ModelDefRGB(const ModelDefRGB&) = delete;
    ModelDefRGB& operator=(const ModelDefRGB&) = delete;
    ~ModelDefRGB() { tamcmc_gpu_rgb_destroy(rgb); tamcmc_gpu_destroy(ctx); }

    double* params_row(int m) { return params.data() + (size_t)m * Nparams; }
    int n_params() const { return Nparams; }
    int n_models() const { return Nmodels; }

    // the constructor's initial models (model_def.cpp:142-153): every chain, no prior short-circuit
    int initialise()
    {
        std::fill(active.begin(), active.end(), (unsigned char)1);
        const int rc = evaluate();
        init_logLikelihood = logLikelihood;
        for (int m = 0; m < Nmodels; m++) logPosterior[(size_t)m] = logLikelihood[(size_t)m] + logPrior[(size_t)m];
        initialised = true;
        return rc;
    }

    // generate_model for ALL chains (model_def.cpp:466-482): chains whose logPrior is -inf are not evaluated and take
    // init_logLikelihood / logPosterior = -inf; a chain whose set-up fails where the reference exits gets NaN (rejected by the caller,
    // MALA.cpp:490,522)
    int generate_models()
    {
        if (!initialised) initialise();
        for (int m = 0; m < Nmodels; m++) active[(size_t)m] = (logPrior[(size_t)m] != -std::numeric_limits<double>::infinity()) ? 1 : 0;
        const std::vector<unsigned char> wanted = active;
        const int rc = evaluate();
        for (int m = 0; m < Nmodels; m++) {
            if (wanted[(size_t)m]) logPosterior[(size_t)m] = logLikelihood[(size_t)m] + logPrior[(size_t)m];
            else { logLikelihood[(size_t)m] = init_logLikelihood[(size_t)m]; logPosterior[(size_t)m] = -std::numeric_limits<double>::infinity(); }
        }
        return rc;
    }

    // call_model_explicit (model_def.cpp:209-218): the model spectrum of one parameter vector (host set-up: one chain, not the hot path)
    std::vector<double> call_model_explicit(const std::vector<double>& params0)
    {
        if ((int)params0.size() < Nparams) throw tamcmc_error(TAMCMC_ERR_ARG, "params0 shorter than Nparams");
        std::vector<double> row((size_t)row_stride, 0.0), m((size_t)Nx);
        int nm = 0;
        check(tamcmc_host_expand_rgb_v4(model_id, params0.data(), plength.data(), step, capacity, row.data(), &nm), "tamcmc_host_expand_rgb_v4");
        check(tamcmc_gpu_model(ctx, 0, row.data(), m.data()), "tamcmc_gpu_model");
        return m;
    }

    tamcmc_gpu_ctx* handle() { return ctx; }

private:
    int evaluate()
    {
        int rc = tamcmc_gpu_rgb_expand(rgb, model_id, params.data(), Nparams, plength.data(), step, Nmodels, capacity, rows, row_stride,
                                       nmodes.data(), expand_status.data(), expand_path.data());
        if (rc != TAMCMC_OK) throw tamcmc_error(rc, std::string("tamcmc_gpu_rgb_expand: ") + tamcmc_gpu_rgb_last_error());
        for (int m = 0; m < Nmodels; m++) if (expand_status[(size_t)m] != TAMCMC_OK) active[(size_t)m] = 0;       // no row for that chain
        rc = tamcmc_gpu_eval(ctx, rows, active.data(), logLikelihood.data(), status.data());
        if (rc != TAMCMC_OK && rc != TAMCMC_ERR_WINDOW && rc != TAMCMC_ERR_NONFINITE) check(rc, "tamcmc_gpu_eval");
        for (int m = 0; m < Nmodels; m++)
            if (expand_status[(size_t)m] != TAMCMC_OK) { logLikelihood[(size_t)m] = std::numeric_limits<double>::quiet_NaN(); rc = TAMCMC_ERR_NONFINITE; }
        return rc;
    }

    tamcmc_gpu_ctx* ctx = nullptr;
    tamcmc_gpu_rgb* rgb = nullptr;
    int model_id;
    std::vector<int> plength;
    int Nmodels, capacity, Nparams = 0, row_stride = 0;
    long Nx;
    double step = 0.0;
    double* rows = nullptr;
    std::vector<unsigned char> active;
    bool initialised = false;
};

}  // namespace tamcmc
