"""Quick kernel-only timing of the C2 workload (used while tuning; not the bench)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g
pkg = g.load_package(); synth = pkg.synth
import bench
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
asym = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
trunc = float(sys.argv[3]) if len(sys.argv) > 3 else 30.0
zeroH = int(sys.argv[4]) if len(sys.argv) > 4 else 0
stars, Ps = [], []
T = synth.tcoefs(10, 1.7)
for s in range(S):
    rng = np.random.default_rng(12345 + s)
    params, pl = synth.classic_params(rng, asym=asym, trunc_c=trunc)
    x = synth.freq_axis(bench.NBINS, 500.0)
    if zeroH: params[:20] = 0.0
    with pkg.Context(pkg.Star(3, pl, len(params), x, np.ones_like(x)), 1, [1.0]) as c0:
        M = c0.model(params)
    y = synth.chi2_2dof_spectrum(rng, M)
    stars.append(pkg.Star(3, pl, len(params), x, y)); Ps.append(synth.perturb_chains(rng, params, pl, 10))
ctx = pkg.Context(stars, 10, T)
P = ctx.pack_params(Ps)
for _ in range(5): ctx.eval(P)
ctx.set_profiling(True)
for _ in range(30): ctx.eval(P)
n, a, b = ctx.kernel_ms()
pairs = ctx.pairs_last()
F = 6.0 * pairs + 20.0 * bench.NBINS * S * 10
print("stars %d asym %g: expand %.1f us  whittle %.1f us  pairs %.3e  alg TF/s %.2f" % (S, asym, 1e3 * a / n, 1e3 * b / n, pairs, F / (b / n * 1e-3) / 1e12))
