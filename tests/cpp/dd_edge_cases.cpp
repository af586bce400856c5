#include "dd_math.cuh"
#include <cstdio>
#include <cstdlib>
#include <random>
int main() {
    std::mt19937_64 rng(99);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    // near multiples of pi/2 (poles and zeros of tan), tiny, large
    for (int k = -2000; k <= 2000; k += 7) {
        for (int r = 0; r < 6; r++) {
            const double d = U(rng) * std::pow(10.0, -2.0 * r - 1);
            const double x = k * 1.5707963267948966 + d;
            printf("tan %a %a\n", x, tamcmc_dd::tan_cr(x));
        }
    }
    for (int e = -300; e <= 6; e += 3) { const double x = U(rng) * std::pow(10.0, e); printf("tan %a %a\n", x, tamcmc_dd::tan_cr(x)); }
    for (int i = 0; i < 2000; i++) { const double x = U(rng) * 3.9e6; printf("tan %a %a\n", x, tamcmc_dd::tan_cr(x)); }
    for (int e = -300; e <= 280; e += 2) { const double t = (U(rng) > 0 ? 1 : -1) * (1.0 + 0.5 * U(rng)) * std::pow(10.0, e); printf("atan %a %a\n", t, tamcmc_dd::atan_cr(t)); }
    printf("tan %a %a\n", 0.0, tamcmc_dd::tan_cr(0.0));
    printf("atan %a %a\n", 0.0, tamcmc_dd::atan_cr(0.0));
    return 0;
}
