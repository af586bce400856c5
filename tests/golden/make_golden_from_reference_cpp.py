"""Writes tests/golden/reference_cpp_vectors.npz from oracle/_ref/libtamcmc_refshim.so, i.e. from the REFERENCE's own
C++ sources compiled against the Eigen-API shim (`make -C oracle ref`).  Run in the build container; the .npz is
committed.  Small cases only (a few thousand bins each)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _cases  # noqa: E402
import _refshim  # noqa: E402
import __graft_entry__ as g  # noqa: E402

synth = g.load_package().synth
R = _refshim.get()
out = {}
k = 0
for mid in _cases.ALL_MODELS:
    for seed in (10, 11):
        params, pl, x = _cases.ms_case(synth, mid, seed=seed, N=3000, asym=(0.0 if seed % 2 == 0 else -23.0), do_amp=seed % 2,
                                       step=synth.RESOL_4YR * 12)
        rc, M = R.call_model(mid, params, pl, x)
        assert rc == 0
        out["model_id_%d" % k] = np.int32(mid)
        out["params_%d" % k] = params
        out["plength_%d" % k] = pl
        out["x_%d" % k] = x
        out["model_%d" % k] = M
        k += 1
out["ncases"] = np.int32(k)
rng = np.random.default_rng(77)
wx = synth.freq_axis(40000, 120.0, synth.RESOL_4YR * 5)
win_in, win_out = [], []
while len(win_in) < 400:
    l = int(rng.integers(0, 4)); fc = rng.uniform(wx[0] + 5, wx[-1] - 5)
    gam = rng.choice([rng.uniform(0.05, 1.0), 1.0, rng.uniform(1.0, 9.0)]); fs = rng.choice([rng.uniform(-1, 1), 1.0, rng.uniform(1, 3)])
    c = rng.choice([10.0, 30.0, 50.0])
    win_in.append([l, fc, gam, fs, c]); win_out.append(R.set_imin_imax(wx, l, fc, gam, fs, c, wx[1] - wx[0]))
out["win_x"] = wx; out["win_in"] = np.array(win_in); out["win_out"] = np.array(win_out, dtype=np.int32)
np.savez_compressed(os.path.join(HERE, "reference_cpp_vectors.npz"), **out)
print("wrote", k, "model cases and", len(win_in), "windows")
