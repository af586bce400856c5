for r in 5 4.5 4; do
  out=$(TAMCMC_GPU_FAR_RATIO=$r timeout 120 python bench.py --steps 400 --warmup 10 --no-cpu-baseline --no-extra 2>/dev/null | tail -1)
  echo "ratio $r $(echo $out | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(round(1e3*j['ms_per_step'],2), round(j['value']), 'fused', round(1e3*j['roofline']['kernel_ms'],2))")"
done
python - <<'PY'
import os, sys, json
sys.path.insert(0, "/root/repo/profiles"); sys.path.insert(0, "/root/repo")
import numpy as np
import __graft_entry__ as g
pkg = g.load_package(); synth = pkg.synth
import bench
def run(asym, ratio):
    os.environ["TAMCMC_GPU_FAR_RATIO"] = str(ratio)
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
for asym in (0.0, 10.0):
    M0, L0 = run(asym, 0)
    for ratio in (5, 4.5, 4):
        M1, L1 = run(asym, ratio)
        print("asym", asym, "ratio", ratio, "model_max_rel %.2e logL_max_rel %.2e" % (float(np.max(np.abs(M1 - M0) / np.abs(M0))), float(np.max(np.abs(L1 - L0) / np.abs(L0)))))
PY
