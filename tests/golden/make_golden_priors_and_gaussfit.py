"""Golden vectors for (1) the generic priors and the prior functions of the Gaussian-envelope models, from the REFERENCE's
own stats_dictionary.cpp / priors_calc.cpp, and (2) the reference's real fixture test/inputs/10280410_Gaussfit.{model,data}
(model_Harvey_Gaussian): its model spectrum and likelihood_chi22p at the .model's initial values, from the REFERENCE's own
functions -- all through oracle/_ref/libtamcmc_refshim.so.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_priors_and_gaussfit.py
Outputs: tests/golden/reference_priors.json, tests/golden/reference_gaussfit_10280410.npz (with copies of the two small
input files' CONTENT as arrays/strings, so the GPU box needs nothing from /root/reference)
"""
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _refshim  # noqa: E402

formats = importlib.import_module("tamcmc-c_b200.formats")
REF = "/root/reference/test/inputs"
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def main():
    R = _refshim.get()
    L = R.L
    L.ref_logP.restype = C.c_double
    L.ref_logP.argtypes = [C.c_int] + [C.c_double] * 5
    L.ref_priors.restype = C.c_double
    L.ref_priors.argtypes = [C.c_int, _dp, C.c_int, _dp, _ip]
    rng = np.random.default_rng(4242)
    prim = []
    for kind in (1, 2, 4, 5, 6, 7, 8, 9, 10):
        for _ in range(12):
            a = float(rng.uniform(-5, 5)); b = a + float(rng.uniform(0.1, 20))
            if kind == 2:
                a, b = float(rng.uniform(-50, 50)), float(rng.uniform(0.01, 30))           # mean, sigma
            if kind in (4, 10):
                a, b = float(rng.uniform(0.01, 5)), float(rng.uniform(5, 500))             # hmin, hmax
            if kind == 9:
                a, b = float(rng.uniform(-1, 0)), float(rng.uniform(0, 1))
            c, d = float(rng.uniform(0.05, 10)), float(rng.uniform(0.05, 10))
            x = float(rng.uniform(a - 3 * (b - a), b + 3 * (b - a))) if kind not in (2, 9) else float(rng.uniform(-100, 100))
            prim.append([kind, a, b, c, d, x, L.ref_logP(kind, a, b, c, d, x)])
        # boundary values
        prim.append([kind, 1.0, 3.0, 0.5, 0.7, 1.0, L.ref_logP(kind, 1.0, 3.0, 0.5, 0.7, 1.0)])
        prim.append([kind, 1.0, 3.0, 0.5, 0.7, 3.0, L.ref_logP(kind, 1.0, 3.0, 0.5, 0.7, 3.0)])
        prim.append([kind, 1.0, 3.0, 0.5, 0.7, 0.0, L.ref_logP(kind, 1.0, 3.0, 0.5, 0.7, 0.0)])

    def call(which, params, kinds, pri):
        p = np.ascontiguousarray(params, dtype=np.float64); k = np.ascontiguousarray(kinds, dtype=np.int32)
        q = np.ascontiguousarray(pri, dtype=np.float64)
        return L.ref_priors(which, p.ctypes.data_as(_dp), len(p), q.ctypes.data_as(_dp), k.ctypes.data_as(_ip))

    # the real fixture: 10 parameters, model_Harvey_Gaussian
    m = formats.read_simple_matrix_model(os.path.join(REF, "10280410_Gaussfit.model"))
    x, y = formats.read_data(os.path.join(REF, "10280410_Gaussfit.data"), xrange=m["xrange"])
    sets = []
    for t in range(40):
        p = m["inputs"] * (1 + (0.0 if t == 0 else 0.25) * rng.standard_normal(10))
        if t % 7 == 3:
            p[9] = 0.5                        # envelope narrower than Dnu/2: rejected (priors_calc.cpp:640-643)
        sets.append(dict(which=1, params=p.tolist(), kinds=m["prior_kinds"].tolist(), pri=m["priors"].tolist(),
                         value=call(1, p, m["prior_kinds"], m["priors"])))
        sets.append(dict(which=-1, params=p.tolist(), kinds=m["prior_kinds"].tolist(), pri=m["priors"].tolist(),
                         value=call(-1, p, m["prior_kinds"], m["priors"])))
    # Kallinger2014 + Gaussian: 19 parameters with a mix of prior kinds
    synth = importlib.import_module("tamcmc-c_b200.synth")
    base = np.concatenate([synth.kallinger_gaussian_params(), [0.5]])          # + omega_numax
    kinds = np.array([4, 2, 1, 2, 0, 4, 4, 1, 2, 0, 1, 2, 0, 1, 4, 1, 7, 13, 0], dtype=np.int32)
    pri = np.full((4, 19), -9999.0)
    for i, k in enumerate(kinds):
        v = base[i]
        if k == 1: pri[0, i], pri[1, i] = v - abs(v), v + abs(v)
        if k == 2: pri[0, i], pri[1, i] = v * 1.01, 0.1 * abs(v) + 0.01
        if k == 4: pri[0, i], pri[1, i] = 0.01 * abs(v), 20 * abs(v)
        if k == 7: pri[0, i], pri[1, i], pri[2, i], pri[3, i] = 0.8 * v, 1.2 * v, 0.1 * v, 0.2 * v
    for t in range(40):
        p = base * (1 + (0.0 if t == 0 else 0.2) * rng.standard_normal(19))
        if t % 9 == 4: p[5] = -abs(p[5])
        if t % 9 == 6: p[17] = -2 * p[15]
        sets.append(dict(which=0, params=p.tolist(), kinds=kinds.tolist(), pri=pri.tolist(), value=call(0, p, kinds, pri)))
    json.dump(dict(primitive=prim, sets=sets), open(os.path.join(HERE, "reference_priors.json"), "w"))

    # the fixture's spectrum and likelihood at the initial values, and at a few perturbed vectors
    rows = np.stack([m["inputs"]] + [m["inputs"] * (1 + 0.05 * rng.standard_normal(10)) for _ in range(1)])
    os.chdir("/tmp")
    models = []
    for r in rows:
        rc, M = R.call_model(1, r, [0] * 11, x)
        assert rc == 0
        models.append(M)
    logL = np.array([R.chi22p(y, M, 1) for M in models])
    np.savez_compressed(os.path.join(HERE, "reference_gaussfit_10280410.npz"), x=x, y=y, rows=rows, model=np.stack(models), logL=logL,
                        model_text=open(os.path.join(REF, "10280410_Gaussfit.model")).read())
    print("fixture: %d bins in [%g, %g], logL(initial) = %.10g, log prior(initial) = %.10g" % (len(x), x[0], x[-1], logL[0], sets[0]["value"]))


if __name__ == "__main__":
    main()
