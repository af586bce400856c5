// CPU-only check of Driver::restore_proposal / restore_variables (host/mcmc_driver.hpp; MALA.cpp:191-246) with an analytic
// Gaussian likelihood: a fresh driver restored from the state of a run holds that state, evaluates the same likelihoods and
// goes on sampling the same target.
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../tamcmc-c_b200/host/mcmc_driver.hpp"
#include "../../tamcmc-c_b200/host/outputs.hpp"

int main(int argc, char** argv)
{
    const int Nparams = 4, Nchains = 3;
    const std::vector<double> centre = {1.0, -2.0, 0.5, 3.0}, width = {0.5, 0.2, 1.0, 0.1};
    tamcmc::DriverConfig cfg;
    cfg.Nchains = Nchains; cfg.Nt_learn = {50, 100, 600}; cfg.seed = 7;
    std::vector<double> T(Nchains);
    for (int m = 0; m < Nchains; m++) T[m] = std::pow(cfg.lambda_temp, m);
    tamcmc::Evaluator ev = [&](const double* p, const unsigned char* act, double* logL) {
        for (int m = 0; m < Nchains; m++) {
            double s = 0.0;
            for (int k = 0; k < Nparams; k++) { const double d = (p[m * Nparams + k] - centre[k]) / width[k]; s += d * d; }
            logL[m] = act[m] ? (-0.5 * s) / T[m] : 0.0;                          // tempered, like model_def.cpp:401
        }
        return 0;
    };
    tamcmc::Prior flat = [](const double*) { return 0.0; };
    const std::vector<int> relax = {0, 1, 3};                                    // parameter 2 stays fixed
    const std::vector<double> err = {0.1, 0.1, 0.1};
    const std::vector<double> p0 = {0.0, 0.0, 0.5, 0.0};
    tamcmc::Driver A(cfg, Nparams, Nparams, p0, relax, err, ev, flat);
    for (long i = 0; i < 800; i++) A.step(i);

    tamcmc::DriverConfig cfgB = cfg; cfgB.seed = 99;
    tamcmc::Driver B(cfgB, Nparams, Nparams, p0, relax, err, ev, flat);
    B.restore_proposal(A.sigma.data(), A.mu.data(), A.covarmat.data());
    B.restore_variables(A.vars.data());
    int bad = 0;
    bad += (B.sigma != A.sigma) + (B.mu != A.mu) + (B.covarmat != A.covarmat) + (B.vars != A.vars);
    for (int m = 0; m < Nchains; m++) {
        bad += (B.logLikelihood[m] != A.logLikelihood[m]) + (B.logPosterior[m] != A.logPosterior[m]);
        bad += (B.params[m * Nparams + 2] != 0.5);                               // the fixed parameter is untouched
        for (int v = 0; v < 3; v++) bad += (B.params[m * Nparams + relax[v]] != A.vars[m * 3 + v]);
    }
    if (bad) { std::printf("restored state differs (%d)\n", bad); return 1; }
    // the restored chain 0 keeps sampling the target: mean within 5 standard errors-ish, acceptance sane
    std::vector<double> sum(3, 0.0);
    const long N = 6000;
    for (long i = 0; i < N; i++) { B.step(100000 + i); for (int v = 0; v < 3; v++) sum[v] += B.vars[v]; }
    for (int v = 0; v < 3; v++) {
        const double mean = sum[v] / N, c = centre[relax[v]], w = width[relax[v]];
        if (!(std::fabs(mean - c) < 0.35 * w)) { std::printf("variable %d: mean %.4f vs %.4f (width %.3f)\n", v, mean, c, w); return 1; }
    }
    const double acc = (double)B.n_accept[0] / N;
    if (!(acc > 0.05 && acc < 0.7)) { std::printf("acceptance %.3f\n", acc); return 1; }
    std::printf("restored: identical state, acceptance(chain 0) %.3f\n", acc);
    if (argc >= 2) {
        // through the reference's restore files (6 significant digits): A's state -> write_restore -> read_restore -> driver C
        namespace out = tamcmc::outputs;
        out::RestoreState st;
        st.Nchains = Nchains; st.Nvars = 3; st.iteration = 800; st.variable_names = {"p0", "p1", "p3"};
        st.vars = st.vars_mean = A.vars; st.sigmas = st.sigmas_mean = A.sigma; st.mus = st.mus_mean = A.mu; st.covarmats = st.covarmats_mean = A.covarmat;
        if (out::write_restore(argv[1], "star", "A", st)) { std::printf("write_restore failed\n"); return 1; }
        out::RestoreState rd;
        if (out::read_restore(argv[1], "star", "A", rd)) { std::printf("read_restore failed\n"); return 1; }
        tamcmc::Driver Cd(cfgB, Nparams, Nparams, p0, relax, err, ev, flat);
        Cd.restore_proposal(rd.sigmas.data(), rd.mus.data(), rd.covarmats.data());
        Cd.restore_variables(rd.vars.data());
        auto close = [](const std::vector<double>& a, const std::vector<double>& b) {
            if (a.size() != b.size()) return false;
            for (size_t i = 0; i < a.size(); i++) if (std::fabs(a[i] - b[i]) > 5e-6 * std::fabs(b[i]) + 1e-300) return false;      // %g keeps 6 digits
            return true;
        };
        if (rd.iteration != 800 || rd.variable_names.size() != 3 || !close(Cd.vars, A.vars) || !close(Cd.sigma, A.sigma) || !close(Cd.mu, A.mu) || !close(Cd.covarmat, A.covarmat)) {
            std::printf("state read back from the restore files differs\n"); return 1;
        }
        for (long i = 0; i < 500; i++) Cd.step(200000 + i);
        for (int m = 0; m < Nchains; m++) if (!std::isfinite(Cd.logLikelihood[m])) { std::printf("non-finite likelihood after the restart\n"); return 1; }
        std::printf("restart through the restore files: ok\n");
    }
    return 0;
}
