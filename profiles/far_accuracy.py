"""Accuracy of the far-field folding on the bench workload (C2, 250k bins, 80 modes, 10 chains): model spectrum and logL of the
default build against the same library with TAMCMC_GPU_FAR_RATIO=0 (every component merged per bin), symmetric and asymmetric
profiles.  Run on a GPU box: python profiles/far_accuracy.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g
pkg = g.load_package(); synth = pkg.synth
import bench

def run(asym, ratio):
    if ratio is None: os.environ.pop("TAMCMC_GPU_FAR_RATIO", None)
    else: os.environ["TAMCMC_GPU_FAR_RATIO"] = str(ratio)
    rng = np.random.default_rng(12345)
    params, pl = synth.classic_params(rng, asym=asym)
    x = synth.freq_axis(bench.NBINS, 500.0)
    with pkg.Context(pkg.Star(3, pl, len(params), x, np.ones_like(x)), 1, [1.0]) as c0:
        M = c0.model(params)
    y = synth.chi2_2dof_spectrum(np.random.default_rng(7), np.maximum(M, 1e-3))
    T = synth.tcoefs(10, 1.7)
    with pkg.Context(pkg.Star(3, pl, len(params), x, y), 10, T) as ctx:
        P = ctx.pack_params([synth.perturb_chains(np.random.default_rng(9), params, pl, 10)])
        logL = np.array(ctx.eval(P)[0]).ravel().copy()
    return M, logL

out = {}
for asym in (0.0, 10.0, -60.0):
    M0, L0 = run(asym, 0)
    for ratio in (None, 6, 12):
        M1, L1 = run(asym, ratio)
        out["asym=%g ratio=%s" % (asym, "default" if ratio is None else ratio)] = {
            "model_max_rel": float(np.max(np.abs(M1 - M0) / np.abs(M0))), "logL_max_rel": float(np.max(np.abs(L1 - L0) / np.abs(L0)))}
print(json.dumps(out, indent=1))
