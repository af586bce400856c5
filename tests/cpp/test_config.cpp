// Drives tamcmc-c_b200/host/config.hpp (tests/test_priors_and_formats.py): test_config <file.cfg> prints the DriverConfig it yields.
#include <cstdio>

#include "../../tamcmc-c_b200/host/config.hpp"

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    tamcmc::config::Groups g;
    const int rc = tamcmc::config::read_cfg(argv[1], g);
    if (rc) { std::printf("rc %d\n", rc); return 1; }
    tamcmc::DriverConfig c;
    const bool ok = tamcmc::config::apply_mala(g, c);
    std::printf("ok %d groups %zu Nchains %d lambda_temp %.17g c0 %.17g epsilon1 %.17g epsi2 %.17g A1 %.17g target_acceptance %.17g dN_mixing %ld Nt_learn",
                (int)ok, g.size(), c.Nchains, c.lambda_temp, c.c0, c.epsilon1, c.epsi2, c.A1, c.target_acceptance, c.dN_mixing);
    for (long v : c.Nt_learn) std::printf(" %ld", v);
    std::printf(" periods_learn");
    for (long v : c.periods_learn) std::printf(" %ld", v);
    std::printf("\n");
    return 0;
}
