"""Full-size golden values for BASELINE config C2 (250 000 bins, 80 modes, 10 chains): the tempered log-likelihoods computed
by the REFERENCE's own model_MS_Global_a1etaa3_HarveyLike_Classic + likelihood_chi22p (oracle/_ref/libtamcmc_refshim.so) for
the seeded synthetic star of bench.py.  Inputs are regenerated from the seed on any machine (numpy + the plain-C oracle for the
noiseless spectrum), only the reference's outputs are stored.  Run in the build container:
    python tests/golden/make_golden_c2_fullsize.py        -> tests/golden/reference_c2_fullsize.json"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _oracle  # noqa: E402
import _refshim  # noqa: E402
import __graft_entry__ as g  # noqa: E402


def c2_inputs(synth, oracle, asym=0.0):
    """The C2 star of bench.py (seed 12345): -> params, plength, x, y, P[10, Nparams], T[10]."""
    rng = np.random.default_rng(12345)
    params, pl = synth.classic_params(rng, asym=asym)
    x = synth.freq_axis(250000, 500.0)
    rc, M = oracle.call_model(3, params, pl, x)
    assert rc == 0
    y = synth.chi2_2dof_spectrum(rng, M)
    P = synth.perturb_chains(rng, params, pl, 10)
    return params, pl, x, y, P, synth.tcoefs(10, 1.7)


def main():
    synth = g.load_package().synth
    O, R = _oracle.get(), _refshim.get()
    out = {}
    for asym in (0.0, 10.0):
        params, pl, x, y, P, T = c2_inputs(synth, O, asym)
        rc, L = R.eval_chains(3, P, pl, x, y, T)
        assert rc == 0 and np.all(np.isfinite(L))
        rc, M0 = R.call_model(3, P[0], pl, x)
        out["asym_%g" % asym] = {"logL_reference": [float(v) for v in L], "y_sum": float(y.sum()), "model0_sum": float(M0.sum()),
                                 "model0_at": {str(i): float(M0[i]) for i in (0, 1234, 77777, 125000, 200001, 249999)}}
        print("asym", asym, L[:3])
    json.dump(out, open(os.path.join(HERE, "reference_c2_fullsize.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
