// The C++ driver (host/mcmc_driver.hpp: adaptive Metropolis + parallel tempering, MALA.cpp:296-553, 623-745) sampling a RED-GIANT fit --
// BASELINE configs C1 / C4, the reference's RGBtests preset on its fixture 10722175 -- with the per-step work on the GPU path:
// every step = ONE tamcmc_gpu_rgb_expand (the asymptotic mixed-mode solve of all chains on the device, rows into the staging block)
// + ONE tamcmc_gpu_eval.  Run twice from the same seed: identical chains.  Optionally the same run with the HOST solver
// (tamcmc_host_expand_rgb_v4 per chain under OpenMP): identical chains again when the device rows equal the host rows, and the speed-up.
//   test_rgb_driver <case.bin> <nsteps> [host]
//   case.bin: header [model_id, N, Nchains, Nparams, capacity, nrelax] + plength[11], x, y, params0, relax index[nrelax], error[nrelax],
//             lo[nrelax], hi[nrelax]; all float64
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../tamcmc-c_b200/host/mcmc_driver.hpp"
#include "../../tamcmc-c_b200/host/model_def_gpu.hpp"

static std::vector<double> rd(FILE* f, size_t n)
{
    std::vector<double> v(n);
    if (fread(v.data(), 8, n, f) != n) { printf("short read\n"); exit(2); }
    return v;
}

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    const long nsteps = atol(argv[2]);
    const bool with_host = argc > 3 && !strcmp(argv[3], "host");
    const std::vector<double> h = rd(f, 17);
    const int model_id = (int)h[0], N = (int)h[1], Nchains = (int)h[2], Nparams = (int)h[3], capacity = (int)h[4], nrelax = (int)h[5];
    int pl[11];
    for (int k = 0; k < 11; k++) pl[k] = (int)h[(size_t)(6 + k)];
    const std::vector<double> x = rd(f, (size_t)N), y = rd(f, (size_t)N), params0 = rd(f, (size_t)Nparams), ridx = rd(f, (size_t)nrelax),
                              err = rd(f, (size_t)nrelax), lo = rd(f, (size_t)nrelax), hi = rd(f, (size_t)nrelax);
    fclose(f);
    std::vector<int> relax((size_t)nrelax);
    for (int v = 0; v < nrelax; v++) relax[(size_t)v] = (int)ridx[(size_t)v];
    const double step = x[2] - x[1];

    tamcmc::DriverConfig cfg;
    cfg.Nchains = Nchains; cfg.lambda_temp = 3.5; cfg.seed = 10722175;                       // config_default.cfg:27,29 of the RGB preset
    cfg.Nt_learn = {100, nsteps / 2, nsteps / 2 + 1}; cfg.periods_learn = {1, 1};
    std::vector<double> Tcoefs((size_t)Nchains);
    for (int m = 0; m < Nchains; m++) Tcoefs[(size_t)m] = std::pow(cfg.lambda_temp, m);

    const int Nnoise = pl[8];
    tamcmc_gpu_star s = tamcmc_gpu_star();
    s.model_id = TAMCMC_MODEL_MODE_TABLE;
    s.plength[0] = capacity; s.plength[1] = 1; s.plength[8] = Nnoise;
    s.Nparams = TAMCMC_MT_HEADER + Nnoise + TAMCMC_MT_STRIDE * capacity;
    s.x = x.data(); s.y = y.data(); s.N = N;
    tamcmc_gpu_ctx* ctx = nullptr;
    int rc = tamcmc_gpu_create(0, 1, &s, Nchains, Tcoefs.data(), 1.0, TAMCMC_LIKELIHOOD_CHI22P, &ctx);
    if (rc) { printf("tamcmc_gpu_create: %s %s\n", tamcmc_gpu_strerror(rc), tamcmc_gpu_last_error()); return 1; }
    tamcmc_gpu_rgb* rgb = nullptr;
    rc = tamcmc_gpu_rgb_create(&rgb, 0, Nchains);
    if (rc) { printf("tamcmc_gpu_rgb_create: %s\n", tamcmc_gpu_rgb_last_error()); return 1; }
    int row_stride = 0;
    double* rows = tamcmc_gpu_params_staging(ctx, &row_stride);
    std::vector<int> est((size_t)Nchains), path((size_t)Nchains), nm((size_t)Nchains);
    std::vector<unsigned char> act((size_t)Nchains);
    long host_handoffs = 0, failed_setups = 0;

    // one MCMC step's likelihood work: set-up of every chain (device), then the batched evaluation
    tamcmc::Evaluator ev_dev = [&](const double* pr, const unsigned char* active, double* L) {
        int r = tamcmc_gpu_rgb_expand(rgb, model_id, pr, Nparams, pl, step, Nchains, capacity, rows, row_stride, nm.data(), est.data(), path.data());
        if (r) return r;
        for (int m = 0; m < Nchains; m++) {
            act[(size_t)m] = (active[m] && est[(size_t)m] == TAMCMC_OK) ? 1 : 0;
            if (est[(size_t)m] != TAMCMC_OK) failed_setups++;
            else if (path[(size_t)m] != 0) host_handoffs++;
        }
        r = tamcmc_gpu_eval(ctx, rows, act.data(), L, nullptr);
        for (int m = 0; m < Nchains; m++) if (!act[(size_t)m]) L[m] = NAN;
        return r;
    };
    tamcmc::Evaluator ev_host = [&](const double* pr, const unsigned char* active, double* L) {
#pragma omp parallel for schedule(dynamic, 1)
        for (int m = 0; m < Nchains; m++) est[(size_t)m] = tamcmc_host_expand_rgb_v4(model_id, pr + (size_t)m * Nparams, pl, step, capacity, rows + (size_t)m * row_stride, &nm[(size_t)m]);
        for (int m = 0; m < Nchains; m++) act[(size_t)m] = (active[m] && est[(size_t)m] == TAMCMC_OK) ? 1 : 0;
        const int r = tamcmc_gpu_eval(ctx, rows, act.data(), L, nullptr);
        for (int m = 0; m < Nchains; m++) if (!act[(size_t)m]) L[m] = NAN;
        return r;
    };
    tamcmc::Prior prior = [&](const double* row) -> double {
        for (int v = 0; v < nrelax; v++) if (row[relax[(size_t)v]] < lo[(size_t)v] || row[relax[(size_t)v]] > hi[(size_t)v]) return -(double)INFINITY;
        return 0.0;
    };

    struct Out { std::vector<double> last, mean; double acc0, swap, secs, L0_first, L0_last; long nanL; };
    auto run = [&](tamcmc::Evaluator ev, long n) {
        tamcmc::Driver d(cfg, Nparams, Nparams, params0, relax, err, ev, prior);
        Out O;
        O.L0_first = d.logLikelihood[0];
        O.mean.assign((size_t)nrelax, 0.0);
        O.nanL = 0;
        const auto t0 = std::chrono::steady_clock::now();
        for (long i = 0; i < n; i++) {
            d.step(i);
            for (int v = 0; v < nrelax; v++) O.mean[(size_t)v] += d.vars[(size_t)v] / (double)n;
            if (std::isnan(d.logLikelihood[0])) O.nanL++;
        }
        O.secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        O.last = d.params;
        O.acc0 = (double)d.n_accept[0] / (double)n;
        O.swap = d.n_swap_tried ? (double)d.n_swap_done / (double)d.n_swap_tried : 0.0;
        O.L0_last = d.logLikelihood[0];
        return O;
    };
    int bad = 0;
    const Out A = run(ev_dev, nsteps), B = run(ev_dev, nsteps);
    if (A.last != B.last) { printf("two runs from the same seed differ\n"); bad++; }
    if (!(std::isfinite(A.L0_first) && std::isfinite(A.L0_last)) || A.nanL) { printf("chain 0: logL %g -> %g, %ld NaN steps\n", A.L0_first, A.L0_last, A.nanL); bad++; }
    if (!(A.acc0 > 0.02 && A.acc0 < 0.95)) { printf("acceptance of chain 0 %.3f\n", A.acc0); bad++; }
    if (!(A.L0_last > A.L0_first - 100.0)) { printf("chain 0 walked away: logL %.3f -> %.3f\n", A.L0_first, A.L0_last); bad++; }
    double host_secs = 0.0;
    bool host_identical = false;
    const long nh = std::min<long>(nsteps, 60);
    if (with_host) {
        const Out Hh = run(ev_host, nh), Dd = run(ev_dev, nh);
        host_secs = Hh.secs;
        host_identical = (Hh.last == Dd.last);
        if (!host_identical) { printf("the chain with the host solver differs from the chain with the device solver after %ld steps\n", nh); bad++; }
    }
    printf("{\"config\": \"red-giant fit with the C++ driver (adaptive Metropolis + parallel tempering), fixture 10722175\", \"chains\": %d, \"bins\": %d, "
           "\"relaxed_variables\": %d, \"steps\": %ld, \"mcmc_steps_per_s\": %.1f, \"evals_per_s\": %.0f, \"ms_per_step\": %.4f, \"acceptance_chain0\": %.3f, "
           "\"swap_rate\": %.3f, \"logL_chain0\": [%.3f, %.3f], \"failed_setups\": %ld, \"host_handoffs\": %ld, \"same_seed_identical\": %s",
           Nchains, N, nrelax, nsteps, nsteps / A.secs, nsteps * (double)Nchains / A.secs, 1e3 * A.secs / nsteps, A.acc0, A.swap, A.L0_first, A.L0_last,
           failed_setups, host_handoffs, (A.last == B.last) ? "true" : "false");
    if (with_host) printf(", \"host_solver\": {\"steps\": %ld, \"mcmc_steps_per_s\": %.2f, \"ms_per_step\": %.3f, \"chain_identical_to_device_solver\": %s}", nh, nh / host_secs,
                          1e3 * host_secs / nh, host_identical ? "true" : "false");
    printf("}\n");
    tamcmc_gpu_rgb_destroy(rgb);
    tamcmc_gpu_destroy(ctx);
    printf(bad ? "FAILED (%d)\n" : "rgb driver: ok\n", bad);
    return bad ? 1 : 0;
}
