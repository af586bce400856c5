// test_mcmc_driver.cpp -- posterior consistency of the fixed-seed adaptive-Metropolis + parallel-tempering driver
// (tamcmc-c_b200/host/mcmc_driver.hpp, a restatement of MALA.cpp:296-553,623-745) when its likelihood comes from
//   (a) the GPU hot path through the C ABI, and
//   (b) the CPU oracle (test infrastructure, linked here only as the checker).
// Same seed, same proposals: the two runs must give the same posterior summaries (BASELINE.json north_star:
// "posterior summaries from a fixed-seed run must be statistically consistent with the reference's").
//
//   test_mcmc_driver <case.bin> <nsteps>          (case format: tests/test_host_cpp.py)
//   test_mcmc_driver <case.bin> <nsteps> batch <nstars>   nstars independent copies of the star (different seeds) driven by
//                                                 BatchDriver: one tamcmc_gpu_eval per step for all stars; checks that star 0
//                                                 reproduces the single-star run bit for bit, prints aggregate star-steps/s
//   test_mcmc_driver <case.bin> <nsteps> bench    GPU likelihood only, every fitted quantity of an MS global fit relaxed
//                                                 (heights, all frequencies, widths, a1, inclination): prints MCMC steps/s
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "../../include/tamcmc_gpu.h"
#include "../../oracle/tamcmc_oracle.h"
#include "../../tamcmc-c_b200/host/mcmc_driver.hpp"

static std::vector<double> rd(FILE* f, size_t n) { std::vector<double> v(n); if (fread(v.data(), 8, n, f) != n) { std::puts("short read"); std::exit(2); } return v; }

struct Summary { std::vector<double> mean, sd; double acc0; double swap; double secs; double secs_sampling; };

int main(int argc, char** argv)
{
    if (argc < 3) { std::puts("usage: test_mcmc_driver case.bin nsteps"); return 2; }
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) { std::perror(argv[1]); return 2; }
    const long nsteps = std::atol(argv[2]);
    const std::vector<double> h = rd(f, 16);
    const int model_id = (int)h[0], Nmodels = (int)h[2], Nparams = (int)h[3];
    const long N = (long)h[1];
    const double p = h[4];
    int pl[11];
    for (int k = 0; k < 11; k++) pl[k] = (int)h[5 + k];
    const std::vector<double> x = rd(f, (size_t)N), y = rd(f, (size_t)N);
    const std::vector<double> T_in = rd(f, (size_t)Nmodels);
    const std::vector<double> P = rd(f, (size_t)Nmodels * Nparams);
    std::fclose(f);
    const std::vector<double> params0(P.begin(), P.begin() + Nparams);
    const bool batch = (argc > 4 && std::string(argv[3]) == "batch");
    const int nbatch = batch ? std::atoi(argv[4]) : 0;
    const bool bench = (argc > 3 && std::string(argv[3]) == "bench") || batch;

    // relaxed variables: all heights, all l=0 frequencies, all widths (the fitted quantities of an MS global fit)
    const int Nmax = pl[0], lmax = pl[1], Nf = pl[2] + pl[3] + pl[4] + pl[5];
    const int o_width = Nmax + lmax + Nf + pl[6];
    std::vector<int> relax;
    std::vector<double> err, lo, hi;
    for (int n = 0; n < Nmax; n++) { relax.push_back(n); err.push_back(0.05 * std::fabs(params0[n])); }
    for (int n = 0; n < (bench ? Nf : pl[2]); n++) { relax.push_back(Nmax + lmax + n); err.push_back(0.05); }
    for (int n = 0; n < pl[7]; n++) { relax.push_back(o_width + n); err.push_back(0.05 * std::fabs(params0[o_width + n])); }
    if (bench) {
        relax.push_back(Nmax + lmax + Nf); err.push_back(0.02);                                   // a1
        relax.push_back(o_width + pl[7] + pl[8]); err.push_back(1.0);                              // inclination
    }
    const size_t nfreq = (size_t)(bench ? Nf : pl[2]);
    for (size_t v = 0; v < relax.size(); v++) {
        const double c = params0[relax[v]];
        double w = (v >= (size_t)Nmax && v < (size_t)Nmax + nfreq) ? 2.0 : 0.6 * std::fabs(c);
        if (bench && v + 1 == relax.size()) w = 40.0;
        lo.push_back(c - w); hi.push_back(c + w);
    }
    tamcmc::Prior prior = [&](const double* row) -> double {     // uniform box: -inf outside (exercises the prior short-circuit, model_def.cpp:469-480)
        for (size_t v = 0; v < relax.size(); v++) if (row[relax[v]] < lo[v] || row[relax[v]] > hi[v]) return -(double)INFINITY;
        return 0.0;
    };

    tamcmc::DriverConfig cfg;
    cfg.Nchains = Nmodels; cfg.lambda_temp = T_in.size() > 1 ? T_in[1] / T_in[0] : 1.7; cfg.seed = 20261018;
    cfg.Nt_learn = {100, nsteps / 2, nsteps / 2 + 1}; cfg.periods_learn = {1, 1};
    std::vector<double> Tcoefs(Nmodels);
    for (int m = 0; m < Nmodels; m++) Tcoefs[m] = std::pow(cfg.lambda_temp, m);

    tamcmc_gpu_star s = {};
    s.model_id = model_id;
    for (int k = 0; k < 11; k++) s.plength[k] = pl[k];
    s.Nparams = Nparams; s.x = x.data(); s.y = y.data(); s.N = N;
    tamcmc_gpu_ctx* ctx = nullptr;
    int rc = tamcmc_gpu_create(0, 1, &s, Nmodels, Tcoefs.data(), p, TAMCMC_LIKELIHOOD_CHI22P, &ctx);
    if (rc) { std::printf("tamcmc_gpu_create: %s %s\n", tamcmc_gpu_strerror(rc), tamcmc_gpu_last_error()); return 1; }
    const int stride = tamcmc_gpu_params_stride(ctx);

    tamcmc::Evaluator ev_gpu = [&](const double* pr, const unsigned char* act, double* L) { return tamcmc_gpu_eval(ctx, pr, act, L, nullptr); };
    tamcmc::Evaluator ev_cpu = [&](const double* pr, const unsigned char* act, double* L) {
        std::vector<double> tmp(Nmodels);
        int r = orc_eval_chains(model_id, pr, stride, pl, x.data(), y.data(), N, Nmodels, Tcoefs.data(), p, tmp.data(), 0);
        for (int m = 0; m < Nmodels; m++) L[m] = act[m] ? tmp[m] : NAN;
        return r;
    };

    // the GPU evaluation in two halves: the driver prepares the next proposal's draws while the kernels run
    tamcmc::AsyncEvaluator async_gpu;
    async_gpu.begin = [&](const double* pr, const unsigned char* act) { return tamcmc_gpu_eval_begin(ctx, pr, act); };
    async_gpu.end = [&](double* L) { return tamcmc_gpu_eval_end(ctx, L, nullptr); };

    auto run = [&](tamcmc::Evaluator ev, bool use_async = false) {
        tamcmc::Driver d(cfg, Nparams, stride, params0, relax, err, ev, prior);
        if (use_async) d.set_async_evaluator(async_gpu);
        const int nv = d.n_vars();
        std::vector<double> s1(nv, 0.0), s2(nv, 0.0);
        long cnt = 0;
        const auto t0 = std::chrono::steady_clock::now();
        auto thalf = t0;
        for (long i = 0; i < nsteps; i++) {
            if (i == nsteps / 2 + 1) thalf = std::chrono::steady_clock::now();        // the learning windows end at nsteps / 2
            d.step(i);
            if (i >= nsteps / 2) { for (int v = 0; v < nv; v++) { const double q = d.vars[v]; s1[v] += q; s2[v] += q * q; } cnt++; }
        }
        Summary S;
        S.secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        S.secs_sampling = std::chrono::duration<double>(std::chrono::steady_clock::now() - thalf).count();
        for (int v = 0; v < nv; v++) { const double m = s1[v] / cnt; S.mean.push_back(m); S.sd.push_back(std::sqrt(std::fmax(s2[v] / cnt - m * m, 0.0))); }
        S.acc0 = (double)d.n_accept[0] / nsteps;
        S.swap = d.n_swap_tried ? (double)d.n_swap_done / d.n_swap_tried : 0.0;
        return S;
    };
    if (batch) {
        // ---- C5-style: nbatch independent stars, one batched evaluation per step ----
        cfg.Nt_learn = {100, nsteps / 2, nsteps / 2 + 1};
        const Summary single = run(ev_gpu);                    // reference: star 0 alone
        tamcmc_gpu_destroy(ctx);
        std::vector<tamcmc_gpu_star> ss((size_t)nbatch, s);
        tamcmc_gpu_ctx* bctx = nullptr;
        rc = tamcmc_gpu_create(0, nbatch, ss.data(), Nmodels, Tcoefs.data(), p, TAMCMC_LIKELIHOOD_CHI22P, &bctx);
        if (rc) { std::printf("tamcmc_gpu_create(batch): %s %s\n", tamcmc_gpu_strerror(rc), tamcmc_gpu_last_error()); return 1; }
        std::vector<std::unique_ptr<tamcmc::Driver>> ds;
        for (int k = 0; k < nbatch; k++) {
            tamcmc::DriverConfig c = cfg;
            c.seed = cfg.seed + (std::uint64_t)k;               // star 0 keeps the seed of the single-star run
            ds.emplace_back(new tamcmc::Driver(c, Nparams, stride, params0, relax, err, tamcmc::Evaluator(), prior, true));
        }
        tamcmc::BatchDriver bd(std::move(ds), [&](const double* pr, const unsigned char* act, double* L) { return tamcmc_gpu_eval(bctx, pr, act, L, nullptr); });
        {
            tamcmc::AsyncEvaluator ab;
            ab.begin = [&](const double* pr, const unsigned char* act) { return tamcmc_gpu_eval_begin(bctx, pr, act); };
            ab.end = [&](double* L) { return tamcmc_gpu_eval_end(bctx, L, nullptr); };
            bd.set_async_evaluator(ab);
        }
        const int nv = bd.stars[0]->n_vars();
        std::vector<double> s1(nv, 0.0);
        long cnt = 0;
        const auto t0 = std::chrono::steady_clock::now();
        for (long i = 0; i < nsteps; i++) {
            bd.step(i);
            if (i >= nsteps / 2) { for (int v = 0; v < nv; v++) s1[v] += bd.stars[0]->vars[v]; cnt++; }
        }
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        tamcmc_gpu_destroy(bctx);
        int bad = 0;
        for (int v = 0; v < nv; v++) if (s1[v] / cnt != single.mean[(size_t)v]) bad++;       // same seed, same data: identical chain
        std::printf("{\"stars\": %d, \"chains\": %d, \"bins\": %ld, \"relaxed_variables\": %d, \"steps\": %ld, \"star_steps_per_s\": %.1f, "
                    "\"evals_per_s\": %.0f, \"ms_per_batched_step\": %.3f, \"star0_matches_single_run\": %s}\n",
                    nbatch, Nmodels, N, nv, nsteps, nbatch * nsteps / secs, nbatch * (double)Nmodels * nsteps / secs, 1e3 * secs / nsteps, bad ? "false" : "true");
        return bad ? 1 : 0;
    }
    if (bench) {
        cfg.Nt_learn = {100, nsteps / 2, nsteps / 2 + 1};
        const Summary G = run(ev_gpu, true);
        tamcmc_gpu_destroy(ctx);
        const long nsamp = nsteps - (nsteps / 2 + 1);
        std::printf("{\"mcmc_steps_per_s\": %.1f, \"evals_per_s\": %.0f, \"mcmc_steps_per_s_sampling_phase\": %.1f, \"mcmc_steps_per_s_learning_phase\": %.1f, "
                    "\"chains\": %d, \"bins\": %ld, \"relaxed_variables\": %zu, \"steps\": %ld, \"acceptance_chain0\": %.3f, \"swap_rate\": %.3f}\n",
                    nsteps / G.secs, nsteps * (double)Nmodels / G.secs, nsamp / G.secs_sampling, (nsteps - nsamp) / (G.secs - G.secs_sampling),
                    Nmodels, N, relax.size(), nsteps, G.acc0, G.swap);
        return 0;
    }
    const Summary G = run(ev_gpu), C = run(ev_cpu), GA = run(ev_gpu, true);
    tamcmc_gpu_destroy(ctx);
    // the two-halves evaluation only moves host work under the kernels: same draws, same chain
    for (size_t v = 0; v < G.mean.size(); v++)
        if (G.mean[v] != GA.mean[v] || G.sd[v] != GA.sd[v]) { std::printf("asynchronous evaluation changed the chain (variable %zu)\n", v); return 1; }
    std::printf("gpu (begin/end): %.1f steps/s, identical chain\n", nsteps / GA.secs);
    std::printf("gpu: %.1f steps/s (%.0f likelihood evals/s), acceptance(chain 0) %.3f, swap rate %.3f\n", nsteps / G.secs, nsteps * Nmodels / G.secs, G.acc0, G.swap);
    std::printf("cpu: %.1f steps/s, acceptance(chain 0) %.3f, swap rate %.3f\n", nsteps / C.secs, C.acc0, C.swap);
    int bad = 0;
    double worst = 0;
    for (size_t v = 0; v < G.mean.size(); v++) {
        const double sd = std::fmax(std::fmax(G.sd[v], C.sd[v]), 1e-12);
        const double dz = std::fabs(G.mean[v] - C.mean[v]) / sd;
        worst = std::fmax(worst, dz);
        if (!(dz < 0.35) || !(G.sd[v] > 0) || !(std::fabs(G.sd[v] - C.sd[v]) < 0.5 * sd)) { std::printf("var %zu: gpu %.6g +- %.3g  cpu %.6g +- %.3g\n", v, G.mean[v], G.sd[v], C.mean[v], C.sd[v]); bad++; }
    }
    if (!(G.acc0 > 0.05 && G.acc0 < 0.8) || std::fabs(G.acc0 - C.acc0) > 0.05) { std::printf("acceptance rates differ or are degenerate\n"); bad++; }
    std::printf("%zu variables, worst |mean_gpu - mean_cpu| / sd = %.3f, failures %d\n", G.mean.size(), worst, bad);
    return bad ? 1 : 0;
}
