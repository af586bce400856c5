"""Generates tests/golden/reference_py_vectors.json by IMPORTING the reference's own Python
restatements of the scalar helpers on the hot path:

    /root/reference/test/lorentzian_test/function_rot.py   (amplitude_ratio, dmm)
    /root/reference/test/lorentzian_test/acoefs.py         (Pslm, eval_acoefs, nunlm_from_acoefs)
    /root/reference/test/lorentzian_test/activity.py       (Qlm, eta0_fct)

Run in the build container only (the reference tree does not travel to the GPU box); the JSON
it writes is committed.  Two shims are needed to import those files with the numpy/matplotlib of
this image: `np.math` (removed in numpy 2) and a dummy `matplotlib.pyplot`.
"""
import json
import math
import os
import sys
import types

import numpy as np

REF = "/root/reference/test/lorentzian_test"

np.math = math  # numpy>=2 dropped np.math; the reference calls np.math.factorial
mpl = types.ModuleType("matplotlib")
plt = types.ModuleType("matplotlib.pyplot")
mpl.pyplot = plt
sys.modules.setdefault("matplotlib", mpl)
sys.modules.setdefault("matplotlib.pyplot", plt)
sys.path.insert(0, REF)

import acoefs as ref_acoefs  # noqa: E402
import activity as ref_activity  # noqa: E402
import function_rot as ref_rot  # noqa: E402

out = {"source": "OthmanB/TAMCMC-C test/lorentzian_test/{function_rot,acoefs,activity}.py", "Pslm": [], "Qlm": [],
       "amplitude_ratio": [], "eval_acoefs": [], "eta0_fct": []}

for s in range(1, 7):
    for l in range(1, 4):
        for m in range(-l, l + 1):
            # the Python version divides by zero where the C++ returns 0 (l=1,s>=3 ; l=2,s>=5 ...)
            try:
                with np.errstate(all="ignore"):
                    v = float(ref_acoefs.Pslm(s, l, m))
            except ZeroDivisionError:
                continue
            if not math.isfinite(v):
                continue
            out["Pslm"].append([s, l, m, repr(v)])

for l in range(1, 4):
    for m in range(-l, l + 1):
        out["Qlm"].append([l, m, repr(float(ref_activity.Qlm(l, m)))])

for l in range(1, 4):
    for inc in (0.0, 5.0, 17.3, 30.0, 45.0, 60.0, 71.9, 85.0, 90.0):
        v = ref_rot.amplitude_ratio(l, inc)
        out["amplitude_ratio"].append([l, inc, [repr(float(t)) for t in v]])

rng = np.random.default_rng(20240229)
for l in range(1, 4):
    for _ in range(4):
        a = [rng.uniform(0.2, 3.0)] + [rng.uniform(-0.1, 0.1) for _ in range(5)]
        if l == 1:
            a[2:] = [0, 0, 0, 0]
        if l == 2:
            a[4:] = [0, 0]
        nu0 = rng.uniform(500, 3500)
        nus = ref_acoefs.nunlm_from_acoefs(nu0, l, a1=a[0], a2=a[1], a3=a[2], a4=a[3], a5=a[4], a6=a[5])
        aj = ref_acoefs.eval_acoefs(l, nus)
        out["eval_acoefs"].append([l, repr(float(nu0)), [repr(float(t)) for t in a], [repr(float(t)) for t in nus],
                                   [repr(float(t)) for t in aj]])

for dnu in (10.0, 55.5, 85.0, 103.2, 135.1, 170.0):
    out["eta0_fct"].append([dnu, repr(float(ref_activity.eta0_fct(Dnu=dnu)))])

dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_py_vectors.json")
with open(dst, "w") as f:
    json.dump(out, f, indent=1)
print("wrote", dst, {k: len(v) for k, v in out.items() if isinstance(v, list)})
