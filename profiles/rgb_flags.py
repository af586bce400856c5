import sys, importlib, numpy as np, collections
sys.path.insert(0, "/root/repo")
import __graft_entry__ as g
pkg = g.load_package()
G = np.load("/root/repo/tests/golden/reference_rgb_vectors.npz")
step = G["x"][2] - G["x"][1]
pl = G["plength0"]
rng = np.random.default_rng(7)
Nmax, lmax, Nfl0 = int(pl[0]), int(pl[1]), int(pl[2])
o = Nmax + lmax + Nfl0
cnt = collections.Counter(); tot = 0
with pkg.RgbExpander(25, pl, step, 140, 20) as rx:
    for call in range(100):
        base = G["params%d" % (call % 4)]
        P = np.tile(base, (20, 1))
        P[:, o] += rng.normal(size=20) * 0.05
        P[:, o + 1] += rng.normal(size=20) * 0.5
        P[:, o + 2] = np.abs(P[:, o + 2] + rng.normal(size=20) * 0.1)
        P[:, o + 3] *= np.exp(rng.normal(size=20) * 0.3)
        P[:, Nmax + lmax:o] += rng.normal(size=(20, Nfl0)) * 0.05
        rows, nm, st, path = rx.expand(P)
        for i in range(20):
            tot += 1
            if st[i] != 0: cnt["status%d" % st[i]] += 1
            elif path[i] != 0: cnt["flag%d" % path[i]] += 1
print("chains", tot, dict(cnt))
