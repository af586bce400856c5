// test_model_def_gpu.cpp -- exercises tamcmc-c_b200/host/model_def_gpu.hpp (the C++ mirror of the reference's
// Model_def hot-path half) exactly like the reference's own call sites do (MALA.cpp:488, model_def.cpp:466-482).
//
//   test_model_def_gpu nogpu            : no CUDA device -> the constructor must throw tamcmc_error(TAMCMC_ERR_CUDA)
//                                         (the product has no CPU fallback); exit 0 if it does
//   test_model_def_gpu <case.bin>       : reads a case written by tests/test_host_cpp.py
//                                         (model id, plength, x, y, Tcoefs, params of Nmodels chains, logPrior,
//                                          oracle logL and oracle model of chain 0), runs generate_models() and
//                                          call_model_explicit() on the GPU and checks parity (1e-10 relative).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../tamcmc-c_b200/host/model_def_gpu.hpp"

static std::vector<double> rd(FILE* f, size_t n) { std::vector<double> v(n); if (fread(v.data(), 8, n, f) != n) { std::puts("short read"); std::exit(2); } return v; }

int main(int argc, char** argv)
{
    if (argc < 2) { std::puts("usage: test_model_def_gpu nogpu|case.bin"); return 2; }
    if (!std::strcmp(argv[1], "nogpu")) {
        tamcmc::StarData s;
        s.model_fct_name_switch = TAMCMC_MODEL_MS_GLOBAL_A1ETAA3_HARVEYLIKE_CLASSIC;
        s.plength = {2, 1, 2, 2, 0, 0, 6, 2, 4, 1, 2};
        s.Nparams = 22;
        s.x = {1, 2, 3, 4}; s.y = {1, 1, 1, 1};
        try {
            tamcmc::ModelDefGPU m({s}, 1, {1.0});
        } catch (const tamcmc::tamcmc_error& e) {
            std::printf("threw as expected: status %d: %s\n", e.status, e.what());
            return e.status == TAMCMC_ERR_CUDA ? 0 : 1;
        }
        std::puts("no exception: a CUDA device is present (run the GPU case instead)");
        return 3;
    }
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) { std::perror(argv[1]); return 2; }
    const std::vector<double> h = rd(f, 16);     // model id, N, Nmodels, Nparams, p, plength[11]
    const int model_id = (int)h[0], Nmodels = (int)h[2], Nparams = (int)h[3];
    const long N = (long)h[1];
    const double p = h[4];
    tamcmc::StarData s;
    s.model_fct_name_switch = model_id;
    for (int k = 0; k < 11; k++) s.plength.push_back((int)h[5 + k]);
    s.Nparams = Nparams;
    s.x = rd(f, (size_t)N); s.y = rd(f, (size_t)N);
    const std::vector<double> T = rd(f, (size_t)Nmodels);
    const std::vector<double> P = rd(f, (size_t)Nmodels * Nparams);
    const std::vector<double> logPrior = rd(f, (size_t)Nmodels);
    const std::vector<double> L_ref = rd(f, (size_t)Nmodels);
    const std::vector<double> M_ref = rd(f, (size_t)N);
    std::fclose(f);

    tamcmc::ModelDefGPU md({s}, Nmodels, T, p);
    for (int m = 0; m < Nmodels; m++) {
        std::memcpy(md.params_row(0, m), &P[(size_t)m * Nparams], sizeof(double) * Nparams);
        md.logPrior[m] = logPrior[m];
    }
    const int rc = md.generate_models();
    if (rc != TAMCMC_OK) { std::printf("generate_models rc=%d\n", rc); return 1; }
    int bad = 0;
    for (int m = 0; m < Nmodels; m++) {
        if (std::isinf(logPrior[m])) {
            // prior short-circuit (model_def.cpp:476-480): init_logLikelihood reused, logPosterior = -inf.  init_logLikelihood is
            // what the reference's constructor computes for EVERY chain from the initial parameters "whatever the situation"
            // (model_def.cpp:142-153) -- here the same vectors, so the oracle's value
            const double rel0 = std::fabs(md.init_logLikelihood[m] - L_ref[m]) / std::fabs(L_ref[m]);
            if (md.logLikelihood[m] != md.init_logLikelihood[m] || !(rel0 < 1e-10) || md.logPosterior[m] != -INFINITY) { std::printf("chain %d: short-circuit not honoured\n", m); bad++; }
            continue;
        }
        const double rel = std::fabs(md.logLikelihood[m] - L_ref[m]) / std::fabs(L_ref[m]);
        const double post = md.logLikelihood[m] + logPrior[m];
        if (!(rel < 1e-10) || md.logPosterior[m] != post) { std::printf("chain %d: logL %.17g ref %.17g rel %.3e\n", m, md.logLikelihood[m], L_ref[m], rel); bad++; }
    }
    {   // the two-halves call gives the same numbers; a second begin before end is refused
        const std::vector<double> L1 = md.logLikelihood;
        md.generate_models_begin();
        bool refused = false;
        try { md.generate_models_begin(); } catch (const tamcmc::tamcmc_error& e) { refused = (e.status == TAMCMC_ERR_ARG); }
        const int rc2 = md.generate_models_end();
        if (rc2 != TAMCMC_OK || !refused || md.logLikelihood != L1) { std::printf("begin/end: rc %d refused %d same %d\n", rc2, (int)refused, (int)(md.logLikelihood == L1)); bad++; }
    }
    const std::vector<double> M = md.call_model_explicit(std::vector<double>(P.begin(), P.begin() + Nparams));
    double worst = 0;
    for (long i = 0; i < N; i++) worst = std::fmax(worst, std::fabs(M[i] - M_ref[i]) / std::fabs(M_ref[i]));
    if (!(worst < 1e-10)) { std::printf("model rel err %.3e\n", worst); bad++; }
    std::printf("chains %d  model rel err %.3e  failures %d\n", Nmodels, worst, bad);
    return bad ? 1 : 0;
}
