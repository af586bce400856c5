"""Golden values of the reference's 1-D tabulated prior logP_tabulated (stats_dictionary.cpp:252-291) from the reference's own
sources (oracle/_ref/libtamcmc_refshim.so: `make -C oracle ref`), on the table it ships (…ajAlm_gate_0.priors) and on a seeded
irregular table.  Run in the build container:  python tests/golden/make_golden_tabulated_logp.py"""
import ctypes as C
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
L = C.CDLL(os.path.join(HERE, "..", "..", "oracle", "_ref", "libtamcmc_refshim.so"))
dp = C.POINTER(C.c_double)
L.ref_logP_tabulated.restype = C.c_double
L.ref_logP_tabulated.argtypes = [dp, dp, C.c_int, C.c_double, C.c_int]
tabs = json.load(open(os.path.join(HERE, "reference_tabulated_priors.json")))
rows = [[float(t) for t in l.split()] for l in tabs["0"].splitlines() if l.strip() and l.strip()[0] not in "#!*"]
t0 = np.array(rows)
rng = np.random.default_rng(21)
xs = np.sort(rng.uniform(0.0, 5.0, 12)); ys = rng.uniform(-0.05, 1.0, 12)          # irregular grid, a few negative PDF values
cases = []
for tx, ty in ((t0[:, 0].copy(), t0[:, 1].copy()), (xs, ys)):
    pts = list(np.linspace(tx[0] - 0.1, tx[-1] + 0.1, 23)) + list(tx) + list(rng.uniform(tx[0], tx[-1], 20))
    for x in pts:
        for nrm in (0, 1):
            v = L.ref_logP_tabulated(tx.ctypes.data_as(dp), ty.ctypes.data_as(dp), len(tx), float(x), nrm)
            cases.append({"tab_x": tx.tolist(), "tab_y": ty.tolist(), "x": float(x), "normalise": nrm, "value": v if np.isfinite(v) else repr(float(v))})
json.dump(cases, open(os.path.join(HERE, "reference_tabulated_logp.json"), "w"))
vals = [c["value"] for c in cases]
print(len(cases), sum(isinstance(v, str) for v in vals), min(v for v in vals if not isinstance(v, str)))
