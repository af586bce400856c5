// test_gaussfit.cpp -- the reference's simplest real analysis end to end on the GPU path: Gaussian-envelope fit
// (model_Harvey_Gaussian, models.cpp:5674-5725) of a real spectrum with the priors of its .model file
// (priors_Harvey_Gaussian, priors_calc.cpp:631-647) sampled by the adaptive Metropolis + parallel-tempering driver.
//   likelihood : tamcmc_gpu_eval (all chains in one launch)          include/tamcmc_gpu.h
//   priors     : tamcmc-c_b200/host/priors.hpp
//   sampler    : tamcmc-c_b200/host/mcmc_driver.hpp
//   test_gaussfit <case.bin> <nsteps>     case = [N, Nparams, Nchains, seed] x[N] y[N] inputs[Np] relax[Np] kinds[Np] priors[4][Np] errors[Np]
// Prints one JSON line: posterior mean / sd of every parameter (coldest chain, second half), acceptance, swap rate, steps/s.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/tamcmc_gpu.h"
#include "../../tamcmc-c_b200/host/mcmc_driver.hpp"
#include "../../tamcmc-c_b200/host/priors.hpp"

static std::vector<double> rd(FILE* f, size_t n) { std::vector<double> v(n); if (fread(v.data(), 8, n, f) != n) { std::puts("short read"); std::exit(2); } return v; }

int main(int argc, char** argv)
{
    if (argc < 3) { std::puts("usage: test_gaussfit case.bin nsteps"); return 2; }
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) { std::perror(argv[1]); return 2; }
    const long nsteps = std::atol(argv[2]);
    const std::vector<double> h = rd(f, 4);
    const long N = (long)h[0];
    const int Np = (int)h[1], Nchains = (int)h[2];
    const std::vector<double> x = rd(f, (size_t)N), y = rd(f, (size_t)N), inputs = rd(f, (size_t)Np), relax_d = rd(f, (size_t)Np),
                              kinds_d = rd(f, (size_t)Np), pri = rd(f, (size_t)4 * Np), errors_all = rd(f, (size_t)Np);
    std::fclose(f);
    std::vector<int> kinds((size_t)Np), relax;
    std::vector<double> err;
    for (int i = 0; i < Np; i++) { kinds[(size_t)i] = (int)kinds_d[(size_t)i]; if (relax_d[(size_t)i] != 0) { relax.push_back(i); err.push_back(errors_all[(size_t)i]); } }
    std::vector<double> q[4];
    for (int r = 0; r < 4; r++) q[r].assign(pri.begin() + (size_t)r * Np, pri.begin() + (size_t)(r + 1) * Np);
    const tamcmc::priors::GenericPriors generic(kinds, q[0], q[1], q[2], q[3]);
    if (!generic.valid()) { std::puts("{\"error\": \"prior kind not supported\"}"); return 1; }

    tamcmc::DriverConfig cfg;
    cfg.Nchains = Nchains; cfg.lambda_temp = 1.7; cfg.seed = (std::uint64_t)h[3];
    cfg.Nt_learn = {200, nsteps / 2, nsteps / 2 + 1}; cfg.periods_learn = {1, 1};
    std::vector<double> Tcoefs((size_t)Nchains);
    for (int m = 0; m < Nchains; m++) Tcoefs[(size_t)m] = std::pow(cfg.lambda_temp, m);

    tamcmc_gpu_star star = {};
    star.model_id = TAMCMC_MODEL_HARVEY_GAUSSIAN;
    star.Nparams = Np; star.x = x.data(); star.y = y.data(); star.N = N;
    tamcmc_gpu_ctx* ctx = nullptr;
    int rc = tamcmc_gpu_create(0, 1, &star, Nchains, Tcoefs.data(), 1.0, TAMCMC_LIKELIHOOD_CHI22P, &ctx);
    if (rc != TAMCMC_OK) { std::printf("{\"error\": \"tamcmc_gpu_create: %s %s\"}\n", tamcmc_gpu_strerror(rc), tamcmc_gpu_last_error()); return 1; }
    const int stride = tamcmc_gpu_params_stride(ctx);
    std::vector<int> st((size_t)Nchains);
    tamcmc::Evaluator ev = [&](const double* P, const unsigned char* act, double* L) -> int { return tamcmc_gpu_eval(ctx, P, act, L, st.data()); };
    tamcmc::Prior pr = [&](const double* row) -> double { return (double)tamcmc::priors::priors_Harvey_Gaussian(row, generic); };

    tamcmc::Driver d(cfg, Np, stride, inputs, relax, err, ev, pr);
    const double logpost0 = d.logPosterior[0];
    std::vector<double> s1((size_t)Np, 0.0), s2((size_t)Np, 0.0);
    long nsum = 0;
    const auto t0 = std::chrono::steady_clock::now();
    for (long i = 0; i < nsteps; i++) {
        d.step(i);
        if (i >= nsteps / 2) {
            for (int k = 0; k < Np; k++) { const double v = d.params[(size_t)k]; s1[(size_t)k] += v; s2[(size_t)k] += v * v; }
            nsum++;
        }
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("{\"steps\": %ld, \"steps_per_s\": %.1f, \"acceptance_cold\": %.4f, \"swap_rate\": %.4f, \"logpost_initial\": %.10g, \"logpost_final\": %.10g, \"mean\": [",
                nsteps, nsteps / secs, (double)d.n_accept[0] / nsteps, d.n_swap_tried ? (double)d.n_swap_done / d.n_swap_tried : 0.0, logpost0, d.logPosterior[0]);
    for (int k = 0; k < Np; k++) std::printf("%s%.10g", k ? ", " : "", s1[(size_t)k] / nsum);
    std::printf("], \"sd\": [");
    for (int k = 0; k < Np; k++) { const double m = s1[(size_t)k] / nsum; std::printf("%s%.10g", k ? ", " : "", std::sqrt(std::fmax(s2[(size_t)k] / nsum - m * m, 0.0))); }
    std::printf("], \"final\": [");
    for (int k = 0; k < Np; k++) std::printf("%s%.17g", k ? ", " : "", d.params[(size_t)k]);
    std::printf("]}\n");
    tamcmc_gpu_destroy(ctx);
    return 0;
}
