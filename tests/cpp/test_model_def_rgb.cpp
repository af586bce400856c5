// ModelDefRGB (tamcmc-c_b200/host/model_def_rgb.hpp) through the C ABI: the red-giant models (ids 25 / 27) behind the reference's
// Model_def interface, with the mixed-mode solve of all chains on the device.
//   test_model_def_rgb nogpu          -> the constructor must fail with TAMCMC_ERR_CUDA (no CPU path)
//   test_model_def_rgb <case.bin>     -> header [model_id, N, Nmodels, Nparams, capacity] + plength[11], x, y, Tcoefs, params, logPrior,
//                                        logL_ref[Nmodels] (tempered), model_ref[N] of chain 0; all float64
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../tamcmc-c_b200/host/model_def_rgb.hpp"

static std::vector<double> rd(FILE* f, size_t n)
{
    std::vector<double> v(n);
    if (fread(v.data(), 8, n, f) != n) { printf("short read\n"); exit(2); }
    return v;
}

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    if (std::string(argv[1]) == "nogpu") {
        std::vector<int> pl = {5, 3, 5, 40, 5, 2, 10, 6, 10, 1, 6};
        std::vector<double> x(100), y(100, 1.0);
        for (int i = 0; i < 100; i++) x[(size_t)i] = 50.0 + 0.01 * i;
        try {
            tamcmc::ModelDefRGB M(25, pl, x, y, 2, {1.0, 2.0});
            printf("constructed without a device?!\n");
            return 1;
        } catch (const tamcmc::tamcmc_error& e) {
            printf("status %d: %s\n", e.status, e.what());
            return e.status == TAMCMC_ERR_CUDA ? 0 : 1;
        }
    }
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    const std::vector<double> h = rd(f, 16);
    const int model_id = (int)h[0], N = (int)h[1], Nmodels = (int)h[2], Nparams = (int)h[3], capacity = (int)h[4];
    std::vector<int> pl(11);
    for (int k = 0; k < 11; k++) pl[(size_t)k] = (int)h[(size_t)(5 + k)];
    const std::vector<double> x = rd(f, (size_t)N), y = rd(f, (size_t)N), T = rd(f, (size_t)Nmodels), P = rd(f, (size_t)Nmodels * Nparams),
                              lp = rd(f, (size_t)Nmodels), Lref = rd(f, (size_t)Nmodels), Mref = rd(f, (size_t)N);
    fclose(f);
    tamcmc::ModelDefRGB M(model_id, pl, x, y, Nmodels, T, capacity);
    if (M.n_params() != Nparams) { printf("Nparams %d != %d\n", M.n_params(), Nparams); return 1; }
    M.params = P;
    int rc = M.initialise();
    if (rc != TAMCMC_OK) { printf("initialise rc %d\n", rc); return 1; }
    int bad = 0;
    for (int m = 0; m < Nmodels; m++) {
        const double rel = std::fabs(M.logLikelihood[(size_t)m] - Lref[(size_t)m]) / std::fabs(Lref[(size_t)m]);
        printf("chain %d: logL %.15g ref %.15g rel %.2e, %d modes, path %d\n", m, M.logLikelihood[(size_t)m], Lref[(size_t)m], rel, M.nmodes[(size_t)m], M.expand_path[(size_t)m]);
        if (!(rel < 1e-10) || M.expand_status[(size_t)m] != TAMCMC_OK || M.expand_path[(size_t)m] != 0) bad++;
    }
    // the prior short-circuit (model_def.cpp:476-480) and a chain whose set-up fails where the reference exits (fmin - Dnu < 0, models.cpp:4852-4858)
    M.logPrior = lp;
    const std::vector<double> init = M.init_logLikelihood;
    const int o_fl0 = pl[0] + pl[1];
    std::vector<double> saved(M.params_row(Nmodels - 1), M.params_row(Nmodels - 1) + Nparams);
    for (int k = 0; k < pl[2]; k++) M.params_row(Nmodels - 1)[o_fl0 + k] = 3.0 + 9.0 * k;
    rc = M.generate_models();
    for (int m = 0; m < Nmodels; m++) {
        const bool masked = std::isinf(lp[(size_t)m]);
        if (masked) { if (M.logLikelihood[(size_t)m] != init[(size_t)m] || !(M.logPosterior[(size_t)m] == -INFINITY)) { printf("chain %d: prior short-circuit not honoured\n", m); bad++; } }
        else if (m == Nmodels - 1) { if (!std::isnan(M.logLikelihood[(size_t)m]) || M.expand_status[(size_t)m] != TAMCMC_ERR_NONFINITE || rc != TAMCMC_ERR_NONFINITE) { printf("chain %d: failed set-up not reported (logL %g, status %d, rc %d)\n", m, M.logLikelihood[(size_t)m], M.expand_status[(size_t)m], rc); bad++; } }
        else if (!(std::fabs(M.logLikelihood[(size_t)m] - Lref[(size_t)m]) < 1e-10 * std::fabs(Lref[(size_t)m])) || M.logPosterior[(size_t)m] != M.logLikelihood[(size_t)m] + lp[(size_t)m]) { printf("chain %d: second evaluation differs\n", m); bad++; }
    }
    std::copy(saved.begin(), saved.end(), M.params_row(Nmodels - 1));
    const std::vector<double> spec = M.call_model_explicit(std::vector<double>(P.begin(), P.begin() + Nparams));
    double worst = 0.0;
    for (int i = 0; i < N; i++) worst = std::fmax(worst, std::fabs(spec[(size_t)i] - Mref[(size_t)i]) / std::fabs(Mref[(size_t)i]));
    printf("call_model_explicit: max rel diff to the reference's model %.2e\n", worst);
    if (!(worst < 1e-10)) bad++;
    printf(bad ? "FAILED (%d)\n" : "model_def_rgb: ok\n", bad);
    return bad ? 1 : 0;
}
