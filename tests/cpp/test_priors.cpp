// Drives tamcmc-c_b200/host/priors.hpp from text lines on stdin (tests/test_priors.py):
//   P kind a b c d x                           -> one primitive prior
//   G|H|K n  params[n]  kinds[n]  pri[4][n]    -> apply_generic_priors | priors_Harvey_Gaussian | priors_Kallinger2014_Gaussian
// Prints one value per line with 21 significant digits.
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "../../tamcmc-c_b200/host/priors.hpp"

using namespace tamcmc::priors;

int main()
{
    std::string tag;
    while (std::cin >> tag) {
        long double r = 0;
        if (tag == "T") {
            // T n tab_x[n] tab_y[n] x normalise  -> logP_tabulated
            int n, nrm; double x;
            std::cin >> n;
            std::vector<double> tx(n), ty(n);
            for (auto& v : tx) std::cin >> v;
            for (auto& v : ty) std::cin >> v;
            std::cin >> x >> nrm;
            r = logP_tabulated(tx.data(), ty.data(), n, x, nrm != 0);
        } else if (tag == "P") {
            int kind; double a, b, c, d, x;
            std::cin >> kind >> a >> b >> c >> d >> x;
            GenericPriors g({kind}, {a}, {b}, {c}, {d});
            r = g.apply(&x);
        } else {
            int n; std::cin >> n;
            std::vector<double> p(n), q[4];
            std::vector<int> k(n);
            for (auto& v : p) std::cin >> v;
            for (auto& v : k) std::cin >> v;
            for (int j = 0; j < 4; j++) { q[j].resize(n); for (auto& v : q[j]) std::cin >> v; }
            GenericPriors g(k, q[0], q[1], q[2], q[3]);
            if (!g.valid()) { std::printf("invalid\n"); continue; }
            r = (tag == "G") ? g.apply(p.data()) : (tag == "H") ? priors_Harvey_Gaussian(p.data(), g) : priors_Kallinger2014_Gaussian(p.data(), g);
        }
        std::printf("%.21Lg\n", r);
    }
    return 0;
}
