"""Full-size golden values for BASELINE configs C3 and C5, computed by the REFERENCE's own functions
(oracle/_ref/libtamcmc_refshim.so) on seeded synthetic inputs that any machine regenerates (numpy + the plain-C oracle).

  C3: model_MS_Global_ajAlm_HarveyLike (models.cpp:1411-1746), 33 modes l <= 2 (the kplr003427720 shape), 10^6 bins, 10 chains,
      gate filter, decompose_Alm = 1, Alm from the SHIPPED 1-degree grids through the reference's own interpolation chain.
      On the GPU box the grids are re-made by the product's GridMaker (token-for-token equal files, tests/test_alm_grids.py).
  C5: 64 independent main-sequence stars x 10 chains x 250 000 bins, model_MS_Global_a1etaa3_HarveyLike_Classic
      (models.cpp:1943-2121), per-star seed, large separation ~ U(60, 100) microHz.
Only the reference's OUTPUTS are stored.  Run in the build container:
    python tests/golden/make_golden_c3_c5.py        -> tests/golden/reference_c3_c5_fullsize.json"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF_GRIDS = "/root/reference/external/Alm/data/Alm_grids_CPP/1deg_grids"
C3_BINS, C3_CHAINS = 1000000, 10
C5_STARS, C5_BINS, C5_CHAINS = 64, 250000, 10


def c3_inputs(synth, oracle, alm):
    """`alm(l, m, theta0, delta, filter_code, user)`: the grid interpolation (reference chain here, product on the GPU box)."""
    rng = np.random.default_rng(3427720)
    params, pl = synth.ajalm_params(rng, Nmax=11, lmax=2, f0=2100.0, dnu=103.0, decompose_Alm=1, filter_code=0,
                                    epsilon=5e-3, theta0=50.0, delta=20.0, trunc_c=30.0)
    x = synth.freq_axis(C3_BINS, 100.0)
    rc, M = oracle.call_model(21, params, pl, x, alm=alm)
    assert rc == 0
    y = synth.chi2_2dof_spectrum(rng, M)
    P = synth.perturb_chains(rng, params, pl, C3_CHAINS)
    return params, pl, x, y, P, synth.tcoefs(C3_CHAINS, 1.7)


def c5_star(synth, oracle, s):
    rng = np.random.default_rng(12345 + s)
    dnu = rng.uniform(60.0, 100.0)
    params, pl = synth.classic_params(rng, f0=620.0 + 0.4 * dnu, dnu=dnu, asym=(0.0 if s % 4 else 8.0))
    x = synth.freq_axis(C5_BINS, 500.0)
    rc, M = oracle.call_model(3, params, pl, x)
    assert rc == 0
    y = synth.chi2_2dof_spectrum(rng, M)
    P = synth.perturb_chains(rng, params, pl, C5_CHAINS)
    return params, pl, x, y, P


def main():
    import _oracle
    import _refshim
    import __graft_entry__ as g
    synth = g.load_package().synth
    O, R = _oracle.get(), _refshim.get()
    out = {}
    assert R.alm_grids_load(REF_GRIDS) == 0
    alm = lambda l, m, t0, de, fc, user: R.Alm_interp(l, m, t0, de, fc)
    params, pl, x, y, P, T = c3_inputs(synth, O, alm)
    L = []
    for c in range(C3_CHAINS):
        rc, M = R.call_model(21, P[c], pl, x)
        assert rc == 0
        L.append(R.chi22p(y, M, 1) / T[c])
        if c == 0:
            m0 = {"model0_sum": float(M.sum()), "model0_at": {str(i): float(M[i]) for i in (0, 54321, 500000, 777777, 999999)}}
    out["c3"] = {"logL_reference": L, "y_sum": float(y.sum()), **m0}
    print("c3", L[:3])
    T5 = synth.tcoefs(C5_CHAINS, 1.7)
    L5, ys = [], []
    for s in range(C5_STARS):
        params, pl, x, y, P = c5_star(synth, O, s)
        rc, L = R.eval_chains(3, P, pl, x, y, T5)
        assert rc == 0 and np.all(np.isfinite(L))
        L5.append([float(v) for v in L]); ys.append(float(y.sum()))
    out["c5"] = {"logL_reference": L5, "y_sum": ys}
    print("c5", L5[0][:2], L5[-1][:2])
    json.dump(out, open(os.path.join(HERE, "reference_c3_c5_fullsize.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
