// ratio_dd (csrc/rgb_solver.cuh) against the reference's long double formula (gnu_fct for one frequency, solver_mm.cpp:172-180, and the
// ratio of :421-427) on random inputs: the two may differ by the extended-precision rounding errors only (a few 1e-15 / |p|).
#include "rgb_solver.cuh"
#include <cstdio>
#include <cstdlib>
#include <random>
typedef long double ld;
int main(int argc, char** argv)
{
    std::mt19937_64 rng(5);
    const long N = atol(argv[1]);
    std::uniform_real_distribution<double> U(30.0, 250.0), D(3.0, 18.0), P(60.0, 400.0), Q(0.05, 0.6), O(-0.45, 0.45), G(-30.0, 30.0);
    long bad = 0; double worst = 0.0;
    for (long i = 0; i < N; i++) {
        tamcmc_rgb::Band B = tamcmc_rgb::Band();
        B.Dnu = D(rng); B.DPl = P(rng); B.q = Q(rng);
        B.nu_p = U(rng);
        const double nu_m = B.nu_p + O(rng) * B.Dnu, nu_g = nu_m + G(rng);
        if (!(nu_g > 5.0) || nu_m == B.nu_p) continue;
        const ld pi = 3.141592653589793238L;
        const ld X = pi * (1. / (ld)nu_m - 1. / (ld)nu_g) * 1e6 / (ld)B.DPl;
        const ld g = (ld)B.Dnu * atanl((ld)B.q * tanl(X)) / pi;
        const ld ratio = g / ((ld)nu_m - (ld)B.nu_p);
        const double r = tamcmc_rgb::ratio_dd(B, nu_g, nu_m);
        const double tol = 1e-14 + 4e-14 / std::fabs(nu_m - B.nu_p);
        const double d = std::fabs((double)(ratio - (ld)r));
        if (d / tol > worst) worst = d / tol;
        if (!(d < tol)) { if (bad < 10) printf("bad: nu_m %a nu_g %a ratio %.20Lg dd %.17g\n", nu_m, nu_g, ratio, r); bad++; }
    }
    printf("N %ld bad %ld worst/tol %.3f\n", N, bad, worst);
    return bad != 0;
}
