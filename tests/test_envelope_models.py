"""Gaussian-envelope models (Config/default/models_ctrl.list ids 0 and 1: model_Kallinger2014_Gaussian, models.cpp:5728-5797,
and model_Harvey_Gaussian, models.cpp:5674-5725): no Lorentzians, a background plus a Gaussian bump, the first stage of the
reference's analysis (numax, envelope width).  Pinned on tests/golden/reference_envelope_vectors.npz, the spectra and
likelihood_chi22p values the REFERENCE's own functions returned (tests/golden/make_golden_envelope_from_reference_cpp.py)."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "reference_envelope_vectors.npz")
RTOL = 1e-10      # north_star: model spectrum and logL within 1e-10 relative (FP64)


def _gold():
    g = np.load(GOLD)
    return [dict(mid=int(g["model_id%d" % i]), x=g["x%d" % i], y=g["y%d" % i], rows=g["params%d" % i], M=g["model%d" % i],
                 logL=g["logL%d" % i]) for i in range(int(g["ncases"]))]


def _rel(a, b):
    return float(np.max(np.abs(a - b) / np.abs(b)))


def test_oracle_matches_reference_envelope_models(pkg, oracle):
    for c in _gold():
        for r, M_ref, L_ref in zip(c["rows"], c["M"], c["logL"]):
            rc, M = oracle.call_model(c["mid"], r, pkg.synth.ENVELOPE_PLENGTH, c["x"])
            assert rc == 0
            assert _rel(M, M_ref) < 1e-13
            assert abs(oracle.chi22p(c["y"], M, 1) - L_ref) <= 1e-12 * abs(L_ref)


@pytest.mark.gpu
def test_gpu_envelope_models_match_reference(pkg, oracle):
    for c in _gold():
        rows, x, y = c["rows"], c["x"], c["y"]
        T = pkg.synth.tcoefs(len(rows), 1.7)
        star = pkg.Star(c["mid"], pkg.synth.ENVELOPE_PLENGTH, rows.shape[1], x, y)
        with pkg.Context(star, len(rows), T) as ctx:
            for r, M_ref in zip(rows, c["M"]):
                assert _rel(ctx.model(r), M_ref) < RTOL            # against the REFERENCE's own output
                rcw, wl, _, _ = ctx.windows(r)
                assert rcw == 0 and len(wl) == 0                   # no Lorentzian windows
            L, st = ctx.eval(rows)
            assert (st == 0).all()
            assert _rel(L[0] * T, c["logL"]) < RTOL                # tempered like model_def.cpp:401
            act = np.array([1, 0, 1], dtype=np.uint8)
            L2, st2 = ctx.eval(rows, active=act)                   # prior short-circuit of one chain (model_def.cpp:476-480)
            assert st2[0, 1] == pkg.CHAIN_INACTIVE and L2[0, 0] == L[0, 0] and L2[0, 2] == L[0, 2]


@pytest.mark.gpu
@pytest.mark.parametrize("N", [9000, 250000])
def test_gpu_envelope_models_full_size_and_mixed_batch(pkg, oracle, N):
    """Both tile sizes; one batch that mixes the two envelope models with a Lorentzian model (every star keeps its own model)."""
    import _cases
    rng = np.random.default_rng(5)
    x = np.arange(N) * (283.2 / N)
    Nch = 4
    T = pkg.synth.tcoefs(Nch, 1.7)
    rows0 = np.stack([pkg.synth.kallinger_gaussian_params(rng, numax=90.0, jitter=0.02) for _ in range(Nch)])
    rows1 = np.stack([pkg.synth.harvey_gaussian_params(rng, numax=140.0, jitter=0.02) for _ in range(Nch)])
    p3, pl3, x3 = _cases.ms_case(pkg.synth, 3, 77, N=12000)
    rows3 = np.stack([p3 * (1 + 1e-3 * rng.standard_normal(p3.size) * (np.arange(p3.size) < pl3[0])) for _ in range(Nch)])
    stars, refs = [], []
    for mid, rows, xs, pl in ((0, rows0, x, pkg.synth.ENVELOPE_PLENGTH), (3, rows3, x3, pl3), (1, rows1, x + 1.5, pkg.synth.ENVELOPE_PLENGTH)):
        rc, M0 = oracle.call_model(mid, rows[0], pl, xs)
        assert rc == 0
        y = M0 * rng.exponential(1.0, len(xs))
        stars.append(pkg.Star(mid, pl, rows.shape[1], xs, y))
        rc, L = oracle.eval_chains(mid, rows, pl, xs, y, T)
        assert rc == 0
        refs.append((M0, L))
    with pkg.Context(stars, Nch, T) as ctx:
        L, st = ctx.eval([rows0, rows3, rows1])
        assert (st == 0).all()
        for s, (M0, L_ref) in enumerate(refs):
            assert _rel(L[s], L_ref) < RTOL
            assert _rel(ctx.model([rows0, rows3, rows1][s][0], star=s), M0) < RTOL
        L_again, _ = ctx.eval([rows0, rows3, rows1])
        assert np.array_equal(L, L_again)                          # bitwise reproducible


@pytest.mark.gpu
def test_gpu_envelope_argument_checks(pkg):
    x = np.arange(4000) * 0.05
    y = np.ones_like(x)
    T = pkg.synth.tcoefs(2, 1.7)
    with pytest.raises(pkg.TamcmcError) as e:      # too few parameters for the fixed layout
        pkg.Context(pkg.Star(0, pkg.synth.ENVELOPE_PLENGTH, 17, x, y), 2, T)
    assert e.value.status == pkg.ERR_ARG
    with pytest.raises(pkg.TamcmcError) as e:      # the Kallinger normalisation integrates the whole spectrum: no bin slices
        pkg.Context(pkg.Star.shard(0, pkg.synth.ENVELOPE_PLENGTH, 18, x, y, 0, 2000), 2, T)
    assert e.value.status == pkg.ERR_ARG
    # the Harvey + Gaussian model has no such sum: a slice evaluates like the same bins of the whole spectrum
    r = np.stack([pkg.synth.harvey_gaussian_params(), pkg.synth.harvey_gaussian_params(numax=100.0)])
    with pkg.Context(pkg.Star(1, pkg.synth.ENVELOPE_PLENGTH, 10, x, y), 2, T) as whole, \
            pkg.Context(pkg.Star.shard(1, pkg.synth.ENVELOPE_PLENGTH, 10, x, y, 1000, 3000), 2, T) as part:
        assert np.allclose(part.model(r[0]), whole.model(r[0])[1000:3000], rtol=1e-13, atol=0)


@pytest.mark.gpu
def test_gpu_envelope_integer_slopes(pkg, oracle):
    """Slopes fixed at exactly 4 or 2 (the usual .model choice) take the product path instead of exp(c ln(x/b))."""
    rng = np.random.default_rng(9)
    x = np.arange(20000) * (283.2 / 20000)          # from 0: the first tiles of model 1 are on the exact background path
    T = pkg.synth.tcoefs(3, 1.7)
    for mid in (0, 1):
        rows = []
        for _ in range(3):
            if mid == 0:
                p = pkg.synth.kallinger_gaussian_params(rng, jitter=0.03)
                p[4], p[9], p[12] = 4.0, 4.0, 2.0
            else:
                p = pkg.synth.harvey_gaussian_params(rng, jitter=0.03)
                p[2], p[5] = 4.0, 2.0
            rows.append(p)
        rows = np.stack(rows)
        rc, M0 = oracle.call_model(mid, rows[0], pkg.synth.ENVELOPE_PLENGTH, x)
        y = M0 * rng.exponential(1.0, len(x))
        rc, L_ref = oracle.eval_chains(mid, rows, pkg.synth.ENVELOPE_PLENGTH, x, y, T)
        assert rc == 0
        with pkg.Context(pkg.Star(mid, pkg.synth.ENVELOPE_PLENGTH, rows.shape[1], x, y), 3, T) as ctx:
            assert _rel(ctx.model(rows[0]), M0) < RTOL
            L, st = ctx.eval(rows)
            assert (st == 0).all() and _rel(L[0], L_ref) < RTOL
