#!/usr/bin/env python
"""Statistics of the device red-giant set-up over a wide cloud of proposals around the fixture's four reference parameter vectors:
which flags send chains to the host solver, and how the mixed-mode frequencies / the other columns compare with the host expander's.
usage: python profiles/rgb_flags.py [ncalls of 20 chains]"""
import collections
import json
import sys

import numpy as np

sys.path.insert(0, "/root/repo")
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
G = np.load("/root/repo/tests/golden/reference_rgb_vectors.npz")
step = G["x"][2] - G["x"][1]
pl = G["plength0"]
rng = np.random.default_rng(7)
Nmax, lmax, Nfl0 = int(pl[0]), int(pl[1]), int(pl[2])
o = Nmax + lmax + Nfl0
nn = int(pl[8])
ncalls = int(sys.argv[1]) if len(sys.argv) > 1 else 100
cnt = collections.Counter()
tot = fc_tot = fc_same = 0
worst_ulp = worst_other = 0.0
with pkg.RgbExpander(25, pl, step, 160, 20) as rx:
    for call in range(ncalls):
        base = G["params%d" % (call % 4)]
        P = np.tile(base, (20, 1))
        P[:, o] += rng.normal(size=20) * 0.05
        P[:, o + 1] += rng.normal(size=20) * 0.5
        P[:, o + 2] = np.abs(P[:, o + 2] + rng.normal(size=20) * 0.1)
        P[:, o + 3] *= np.exp(rng.normal(size=20) * 0.3)
        P[:, Nmax + lmax:o] += rng.normal(size=(20, Nfl0)) * 0.05
        P[:, -3] = call % 2                                   # both solver entry points
        rows, nm, st, path = rx.expand(P)
        for i in range(20):
            tot += 1
            if st[i] != 0:
                cnt["status%d" % st[i]] += 1
                continue
            if path[i] != 0:
                cnt["flag%d" % path[i]] += 1
            row, n = pkg.expand_rgb_v4(25, P[i], pl, step, 160)
            assert n == nm[i]
            a, b = row[4 + nn:4 + nn + 20 * n].reshape(n, 20), rows[i, 4 + nn:4 + nn + 20 * n].reshape(n, 20)
            l1 = a[:, 0] == 1
            fc_tot += int(l1.sum())
            fc_same += int((a[l1, 1] == b[l1, 1]).sum())
            worst_ulp = max(worst_ulp, float(np.max(np.abs(a[:, 1] - b[:, 1]) / np.spacing(a[:, 1]))))
            worst_other = max(worst_other, float(np.max(np.abs(a - b) / np.maximum(np.abs(a), 1e-300))))
    n_setups, n_host = rx.counts()
print(json.dumps({"chains": tot, "not_solved_on_device": dict(cnt), "handle_counts": [n_setups, n_host], "mixed_mode_frequencies": fc_tot,
                  "identical_to_host_solver": fc_same, "max_frequency_difference_ulp": worst_ulp, "max_rel_difference_any_column": worst_other}))
