// The LAW the restated sampler draws from (host/mcmc_driver.hpp), not only its determinism -- CPU only, analytic target.
// Target: a correlated 3-D Gaussian N(c, S) as the untempered log-likelihood, flat prior, Nchains tempered chains with
// T_m = lambda^m (MALA.cpp:98-99).  Chain m stores and is judged on logL / T_m (model_def.cpp:401), so it must sample
// N(c, T_m S): checked are
//   * chain 0: mean and the full covariance matrix against S;
//   * every hotter chain m: covariance = T_m * S (the tempering law of update_position_MH, MALA.cpp:463-553);
//   * the acceptance rate the Robbins-Monro scale update steers to (update_proposal, MALA.cpp:296-319): ~ target_acceptance;
//   * parallel tempering (MALA.cpp:397-461): the observed swap rate against the expectation of
//     min(1, exp[(l(x_A) - l(x_B)) (1/T_B - 1/T_A)]) over independent x_A ~ N(c, T_A S), x_B ~ N(c, T_B S), i.e. over two
//     independent chi^2_3 variables (evaluated here by quadrature-grade Monte Carlo with an independent generator), and that
//     swaps leave every chain's law intact (the covariances above are measured WITH swaps on).
// A second run with the fatal-status path checks that a failing evaluator rejects every proposal.
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#include "../../tamcmc-c_b200/host/mcmc_driver.hpp"

int main()
{
    const int d = 3, Nchains = 4;
    const double lambda = 2.0;
    const double c[d] = {1.0, -2.0, 0.5};
    const double S[d][d] = {{1.0, 0.6, -0.2}, {0.6, 2.0, 0.3}, {-0.2, 0.3, 0.5}};
    // precision matrix P = S^-1 (adjugate / determinant)
    double P[d][d];
    {
        const double det = S[0][0] * (S[1][1] * S[2][2] - S[1][2] * S[2][1]) - S[0][1] * (S[1][0] * S[2][2] - S[1][2] * S[2][0])
                         + S[0][2] * (S[1][0] * S[2][1] - S[1][1] * S[2][0]);
        P[0][0] = (S[1][1] * S[2][2] - S[1][2] * S[2][1]) / det; P[0][1] = (S[0][2] * S[2][1] - S[0][1] * S[2][2]) / det; P[0][2] = (S[0][1] * S[1][2] - S[0][2] * S[1][1]) / det;
        P[1][0] = P[0][1]; P[1][1] = (S[0][0] * S[2][2] - S[0][2] * S[2][0]) / det; P[1][2] = (S[0][2] * S[1][0] - S[0][0] * S[1][2]) / det;
        P[2][0] = P[0][2]; P[2][1] = P[1][2]; P[2][2] = (S[0][0] * S[1][1] - S[0][1] * S[1][0]) / det;
    }
    tamcmc::DriverConfig cfg;
    cfg.Nchains = Nchains; cfg.lambda_temp = lambda; cfg.seed = 20261018; cfg.dN_mixing = 1;
    cfg.Nt_learn = {200, 6000, 30000};           // two learning windows, then a frozen proposal (config_default.cfg:17 shape)
    cfg.periods_learn = {1, 1};
    std::vector<double> T(Nchains);
    for (int m = 0; m < Nchains; m++) T[m] = std::pow(lambda, m);
    bool fail_mode = false;
    tamcmc::Evaluator ev = [&](const double* p, const unsigned char* act, double* logL) {
        if (fail_mode) return 3;                                                 // TAMCMC_ERR_CUDA: nothing written
        for (int m = 0; m < Nchains; m++) {
            double q = 0.0;
            for (int a = 0; a < d; a++) for (int b = 0; b < d; b++) q += (p[m * d + a] - c[a]) * P[a][b] * (p[m * d + b] - c[b]);
            logL[m] = act[m] ? (-0.5 * q) / T[m] : 0.0;
        }
        return 0;
    };
    tamcmc::Prior flat = [](const double*) { return 0.0; };
    const std::vector<int> relax = {0, 1, 2};
    const std::vector<double> err = {0.3, 0.3, 0.3}, p0 = {0.0, 0.0, 0.0};
    tamcmc::Driver D(cfg, d, d, p0, relax, err, ev, flat);
    const long Nburn = 40000, N = 600000;
    for (long i = 0; i < Nburn; i++) if (D.step(i) != 0) { std::printf("unexpected status\n"); return 1; }
    std::vector<double> mean((size_t)Nchains * d, 0.0), cov((size_t)Nchains * d * d, 0.0);
    const long swaps0_t = D.n_swap_tried, swaps0_d = D.n_swap_done;
    std::vector<long> acc0 = D.n_accept;
    for (long i = 0; i < N; i++) {
        D.step(Nburn + i);
        for (int m = 0; m < Nchains; m++)
            for (int a = 0; a < d; a++) {
                const double xa = D.vars[(size_t)m * d + a] - c[a];
                mean[(size_t)m * d + a] += xa;
                for (int b = 0; b < d; b++) cov[((size_t)m * d + a) * d + b] += xa * (D.vars[(size_t)m * d + b] - c[b]);
            }
    }
    int bad = 0;
    for (int m = 0; m < Nchains; m++) {
        for (int a = 0; a < d; a++) {
            const double mu = mean[(size_t)m * d + a] / N, sd = std::sqrt(T[m] * S[a][a]);
            if (!(std::fabs(mu) < 0.04 * sd)) { std::printf("chain %d: mean[%d] off by %.4f sd\n", m, a, mu / sd); bad++; }
        }
        for (int a = 0; a < d; a++)
            for (int b = 0; b < d; b++) {
                const double mua = mean[(size_t)m * d + a] / N, mub = mean[(size_t)m * d + b] / N;
                const double cv = cov[((size_t)m * d + a) * d + b] / N - mua * mub, expct = T[m] * S[a][b];
                const double scale = T[m] * std::sqrt(S[a][a] * S[b][b]);
                if (!(std::fabs(cv - expct) < 0.04 * scale)) { std::printf("chain %d: cov[%d][%d] = %.4f, expected %.4f (T = %g)\n", m, a, b, cv, expct, T[m]); bad++; }
            }
        const double acc = (double)(D.n_accept[(size_t)m] - acc0[(size_t)m]) / N;
        if (!(acc > 0.15 && acc < 0.35)) { std::printf("chain %d: acceptance %.3f, target %.3f\n", m, acc, cfg.target_acceptance); bad++; }
        std::printf("chain %d (T = %g): var/T = %.4f %.4f %.4f (S: %.1f %.1f %.1f), acceptance %.3f\n", m, T[m],
                    (cov[((size_t)m * d + 0) * d + 0] / N) / T[m], (cov[((size_t)m * d + 1) * d + 1] / N) / T[m], (cov[((size_t)m * d + 2) * d + 2] / N) / T[m],
                    S[0][0], S[1][1], S[2][2], acc);
    }
    // expected swap acceptance of an adjacent pair: -2 l(x) = T chi^2_3 under chain T  ->  r = exp[-(qA - lambda qB)(1 - lambda) / (2 lambda)] ... sign below
    {
        std::mt19937_64 g(999);
        std::chi_squared_distribution<double> chi(3.0);
        double e = 0.0;
        const long M = 4000000;
        for (long k = 0; k < M; k++) {
            const double qA = chi(g), qB = chi(g);
            // l(x_A) = -T_A qA / 2, l(x_B) = -T_B qB / 2, T_B = lambda T_A: (l_A - l_B)(1/T_B - 1/T_A) = -(qA - lambda qB)(1 - lambda) / (2 lambda)
            const double r = std::exp(-(qA - lambda * qB) * (1.0 - lambda) / (2.0 * lambda));
            e += r < 1.0 ? r : 1.0;
        }
        e /= M;
        const double got = (double)(D.n_swap_done - swaps0_d) / (double)(D.n_swap_tried - swaps0_t);
        std::printf("swap rate %.4f, expected %.4f (%ld attempts)\n", got, e, D.n_swap_tried - swaps0_t);
        if (!(std::fabs(got - e) < 0.01)) bad++;
    }
    // a failing evaluator: every proposal rejected, the state stays where it was, the status reaches the caller
    {
        const std::vector<double> before = D.vars, Lb = D.logLikelihood;
        fail_mode = true;
        tamcmc::DriverConfig nc = cfg;
        for (long i = 0; i < 50; i++) if (D.step(Nburn + N + i) != 3) { std::printf("fatal status not reported\n"); bad++; break; }
        // (swaps of the CURRENT states may still happen: compare as multisets through the sorted first coordinates)
        std::vector<double> a, b;
        for (int m = 0; m < Nchains; m++) { a.push_back(before[(size_t)m * d]); b.push_back(D.vars[(size_t)m * d]); }
        std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
        if (a != b) { std::printf("positions moved on failed evaluations\n"); bad++; }
        if (D.last_eval_status != 3) bad++;
        bool threw = false;
        try { tamcmc::Driver E(nc, d, d, p0, relax, err, ev, flat); } catch (const std::runtime_error&) { threw = true; }
        if (!threw) { std::printf("constructor accepted a failed initial evaluation\n"); bad++; }
    }
    if (bad) { std::printf("FAILED (%d)\n", bad); return 1; }
    std::printf("driver law: ok\n");
    return 0;
}
