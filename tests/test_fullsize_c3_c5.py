"""BASELINE configs C3 (ajAlm, 10^6 bins) and C5 (64 stars x 10 chains x 250k bins) at their stated sizes against the
log-likelihoods the REFERENCE's own functions returned for the same seeded inputs
(tests/golden/reference_c3_c5_fullsize.json, written by tests/golden/make_golden_c3_c5.py in the build container)."""
import importlib.util
import json
import os

import numpy as np
import pytest

RTOL = 1e-10
GDIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _mod():
    spec = importlib.util.spec_from_file_location("make_golden_c3_c5", os.path.join(GDIR, "make_golden_c3_c5.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def gold():
    return json.load(open(os.path.join(GDIR, "reference_c3_c5_fullsize.json")))


@pytest.fixture(scope="module")
def grids(pkg, tmp_path_factory):
    d = str(tmp_path_factory.mktemp("alm_grids_c3"))
    pkg.AlmGrids.make(d, 0)
    pkg.AlmGrids.make(d, 2)
    return pkg.AlmGrids(d)


def test_c3_oracle_against_reference_golden(pkg, oracle, gold, grids):
    """CPU leg: the regenerated inputs are the ones the reference saw, and the oracle (fed the product's grid interpolation of
    the re-made grids) reproduces the reference's log-likelihood of chain 0 on all 10^6 bins."""
    mod = _mod()
    alm = lambda l, m, t0, de, fc, user: grids(l, m, t0, de, fc)
    params, pl, x, y, P, T = mod.c3_inputs(pkg.synth, oracle, alm)
    assert abs(y.sum() - gold["c3"]["y_sum"]) <= 1e-12 * abs(gold["c3"]["y_sum"])
    rc, M = oracle.call_model(21, P[0], pl, x, alm=alm)
    assert rc == 0
    L0 = oracle.call_likelihood(y, M, 1.0, T[0])
    assert abs(L0 - gold["c3"]["logL_reference"][0]) <= 1e-11 * abs(L0)


def test_c5_oracle_against_reference_golden(pkg, oracle, gold):
    mod = _mod()
    T = pkg.synth.tcoefs(mod.C5_CHAINS, 1.7)
    for s in (0, 63):
        params, pl, x, y, P = mod.c5_star(pkg.synth, oracle, s)
        assert abs(y.sum() - gold["c5"]["y_sum"][s]) <= 1e-13 * abs(y.sum())
        rc, L = oracle.eval_chains(3, P[:2], pl, x, y, T[:2])
        assert rc == 0
        Lr = np.array(gold["c5"]["logL_reference"][s][:2])
        assert np.max(np.abs(L - Lr) / np.abs(Lr)) < 1e-12


@pytest.mark.gpu
def test_c3_fullsize_gpu_against_reference_golden(pkg, oracle, gold, grids):
    """C3 at 10^6 bins from a REFERENCE parameter vector: host expander with the grid interpolation (no caller callback) ->
    mode table -> GPU, all 10 chains, against the reference's log-likelihoods."""
    mod = _mod()
    alm = lambda l, m, t0, de, fc, user: grids(l, m, t0, de, fc)
    params, pl, x, y, P, T = mod.c3_inputs(pkg.synth, oracle, alm)
    assert abs(y.sum() - gold["c3"]["y_sum"]) <= 1e-12 * abs(gold["c3"]["y_sum"])
    cap = int(pl[2:6].sum())
    rows = np.stack([pkg.expand_ajAlm(P[c], pl, cap, alm=grids)[0] for c in range(mod.C3_CHAINS)])
    mpl = pkg.synth.mode_table_plength(cap, int(pl[8]), 0)
    with pkg.Context(pkg.Star(pkg.synth.MODEL_MODE_TABLE, mpl, rows.shape[1], x, y), mod.C3_CHAINS, T) as ctx:
        L, st = ctx.eval(rows)
        assert (st == 0).all()
        Lr = np.array(gold["c3"]["logL_reference"])
        assert np.max(np.abs(L[0] - Lr) / np.abs(Lr)) < RTOL
        M0 = ctx.model(rows[0])
        assert abs(M0.sum() - gold["c3"]["model0_sum"]) <= 1e-11 * abs(gold["c3"]["model0_sum"])
        for i, v in gold["c3"]["model0_at"].items():
            assert abs(M0[int(i)] - v) <= RTOL * abs(v)


@pytest.mark.gpu
def test_c5_64_stars_gpu_against_reference_golden(pkg, oracle, gold):
    """C5: 64 independent stars x 10 chains in ONE batched launch against the reference's 640 log-likelihoods."""
    mod = _mod()
    T = pkg.synth.tcoefs(mod.C5_CHAINS, 1.7)
    stars, Ps = [], []
    for s in range(mod.C5_STARS):
        params, pl, x, y, P = mod.c5_star(pkg.synth, oracle, s)
        assert abs(y.sum() - gold["c5"]["y_sum"][s]) <= 1e-13 * abs(y.sum())
        stars.append(pkg.Star(3, pl, len(params), x, y))
        Ps.append(P)
    with pkg.Context(stars, mod.C5_CHAINS, T) as ctx:
        L, st = ctx.eval(ctx.pack_params(Ps))
        assert (st == 0).all()
        Lr = np.array(gold["c5"]["logL_reference"])
        assert L.shape == Lr.shape == (mod.C5_STARS, mod.C5_CHAINS)
        assert np.max(np.abs(L - Lr) / np.abs(Lr)) < RTOL
        L2, _ = ctx.eval(ctx.pack_params(Ps))
        assert np.array_equal(L, L2)                                   # bitwise reproducible
