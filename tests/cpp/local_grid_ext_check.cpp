#include "rgb_solver.cuh"
#include <cstdio>
#include <cstdlib>
#include <random>
typedef long double ld;
int main(int argc, char** argv) {
    std::mt19937_64 rng(7);
    const long N = atol(argv[1]);
    std::uniform_real_distribution<double> U(5.0, 400.0), R(0.003, 0.12);
    const double facts[3] = {0.04, 0.01, 0.005};
    long bad = 0, notok = 0, n799 = 0, n800 = 0;
    for (long i = 0; i < N; i++) {
        const double nu = U(rng), resol_d = (i % 3 == 0) ? 0.007871049999991442 : R(rng), fact = facts[i % 3];
        const ld resol = resol_d, factor = fact;
        const ld range_min = nu - 2 * resol, range_max = nu + 2 * resol;
        const long n_ref = (long)((range_max - range_min) / (resol * factor));
        const double lo_ref = (double)range_min, hi_ref = (double)range_max;
        const ld D = resol * factor;
        const double Dh = (double)D, Dl = (double)(D - (ld)Dh);
        double lo, hi; int n;
        if (!tamcmc_rgb::local_grid_ext(nu, (double)(2 * resol), Dh, Dl, lo, hi, n)) { notok++; continue; }
        if (lo != lo_ref || hi != hi_ref || n != n_ref) { if (bad < 10) printf("bad nu=%a resol=%a fact=%g: n %d ref %ld lo %a %a hi %a %a\n", nu, resol_d, fact, n, n_ref, lo, lo_ref, hi, hi_ref); bad++; }
        if (fact == 0.005) { if (n_ref == 799) n799++; if (n_ref == 800) n800++; }
    }
    printf("N %ld bad %ld notok %ld (799: %ld, 800: %ld)\n", N, bad, notok, n799, n800);
    return bad != 0;
}
