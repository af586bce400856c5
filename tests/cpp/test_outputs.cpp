// Drives tamcmc-c_b200/host/outputs.hpp (tests/test_priors_and_formats.py): writes the header of the reference's own
// Gaussian-envelope run (metadata on the command line side is fixed here) and two buffers of synthetic samples.
//   usage: test_outputs <prefix>
#include <cstdio>
#include <string>
#include <vector>

#include "../../tamcmc-c_b200/host/outputs.hpp"

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    if (argc >= 5 && std::string(argv[1]) == "restore") {
        // test_outputs restore <dir> <star_id> <phase>: prints what read_restore found, 17 significant digits per value
        tamcmc::outputs::RestoreState st;
        const int rc = tamcmc::outputs::read_restore(argv[2], argv[3], argv[4], st);
        std::printf("rc %d Nchains %d Nvars %d iteration %ld names %zu last %s\n", rc, st.Nchains, st.Nvars, st.iteration, st.variable_names.size(),
                    st.variable_names.empty() ? "" : st.variable_names.back().c_str());
        const std::vector<double>* all[8] = {&st.vars, &st.vars_mean, &st.sigmas, &st.sigmas_mean, &st.mus, &st.mus_mean, &st.covarmats, &st.covarmats_mean};
        for (const auto* v : all) { std::printf("%zu", v->size()); for (double x : *v) std::printf(" %.17g", x); std::printf("\n"); }
        if (argc >= 6 && !rc) {
            // ... and writes it back under <outdir> = argv[5] (round trip of write_restore)
            if (tamcmc::outputs::write_restore(argv[5], argv[3], argv[4], st)) return 1;
        }
        return rc ? 1 : 0;
    }
    tamcmc::outputs::ParamsMeta m;
    m.Nsamples = 100000; m.Nchains = 4;
    m.relax = {1, 1, 0, 1, 1, 1, 1, 1, 1, 1};
    m.plength = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
    m.cons_names = {"p1"}; m.cons_values = {4.0};
    m.var_names = {"H1", "tc1", "H2", "tc2", "p2", "B0", "Amax", "numax", "Gauss_sigma"};
    const int nv = 9;
    auto value = [&](long i, int c, int v) { return 1000.0 * i + 10.0 * c + v + 0.125; };
    for (int part = 0; part < 2; part++) {
        const long n0 = part ? 6 : 0, n = part ? 4 : 6;
        std::vector<double> buf((size_t)n * m.Nchains * nv);
        for (long i = 0; i < n; i++) for (int c = 0; c < m.Nchains; c++) for (int v = 0; v < nv; v++) buf[((size_t)i * m.Nchains + c) * nv + v] = value(n0 + i, c, v);
        const int rc = tamcmc::outputs::write_params(argv[1], m, buf.data(), n, 99999, part == 0);
        if (rc) { std::printf("write_params rc=%d\n", rc); return 1; }
    }
    std::printf("%s", tamcmc::outputs::eigen_row(std::vector<double>{0.5, 12.25}, false).c_str());
    return 0;
}
