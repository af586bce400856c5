"""model_MS_Global_ajAlm_HarveyLike (BASELINE config C3; models.cpp:1411-1746) through the host expander
tamcmc_host_expand_ajAlm -> mode table -> GPU.  The oracle's restatement of that model function takes the Alm value
from a callback (the reference interpolates GSL grids there: un-vendored, parity unpinned at that boundary --
SURVEY.md 8c); both sides are fed the same Alm here."""
import numpy as np
import pytest

RTOL = 1e-10


def test_host_alm_sum_rule_and_quadrature(pkg):
    """sum_m |Y_lm|^2 = (2l+1)/4pi  =>  sum_m Alm(l,m) = (2l+1)(cos(tmin) - cos(tmax)) for the gate filter; each value
    against an independent scipy quadrature of 4 pi |Y_lm|^2 sin(theta)."""
    from scipy import integrate, special
    rng = np.random.default_rng(0)
    for _ in range(10):
        th0, de = rng.uniform(0.1, 1.4), rng.uniform(0.05, 0.8)
        tmin, tmax = max(th0 - de / 2, 0.0), min(th0 + de / 2, np.pi / 2)
        for l in (1, 2, 3):
            vals = [pkg.host_alm(l, m, th0, de, 0) for m in range(-l, l + 1)]
            assert np.isclose(sum(vals), (2 * l + 1) * (np.cos(tmin) - np.cos(tmax)), rtol=1e-13)
            for m in range(-l, l + 1):
                f = lambda t: 4 * np.pi * abs(special.sph_harm_y(l, m, t, 0.0)) ** 2 * np.sin(t)
                ref = integrate.quad(f, tmin, tmax, epsabs=1e-14, epsrel=1e-13)[0]
                assert np.isclose(vals[m + l], ref, rtol=1e-11)
            assert vals == vals[::-1]                      # Alm(l, m) = Alm(l, -m)
    assert pkg.host_alm(2, 1, 0.7, 0.0, 0) == 0.0          # delta == 0 (activity.cpp:196-198)
    assert pkg.host_alm(2, 3, 0.7, 0.1, 0) == -10.0        # |m| > l (activity.cpp:240-243)
    # triangle filter: positive, below the gate of the same band, same hemisphere doubling
    for l in (1, 2, 3):
        for m in range(0, l + 1):
            t, g = pkg.host_alm(l, m, 0.9, 0.4, 2), pkg.host_alm(l, m, 0.9, 0.4, 0)
            assert 0 < t < g


@pytest.mark.parametrize("decompose,filter_code", [(-1, 0), (0, 0), (1, 0), (2, 0), (1, 2), (-1, 2)])
def test_host_expander_matches_oracle_model(pkg, oracle, decompose, filter_code):
    rng = np.random.default_rng(10 + decompose)
    asym = 0.0 if decompose % 2 else 25.0
    params, pl = pkg.synth.ajalm_params(rng, Nmax=6, lmax=3, f0=1000.0, dnu=70.0, decompose_Alm=decompose, filter_code=filter_code,
                                        asym=asym, do_amp=int(decompose == 0), trunc_c=20.0)
    x = pkg.synth.freq_axis(25000, 950.0, 0.02)
    alm = lambda l, m, t0, de, fc, user: pkg.host_alm(l, m, t0, de, fc)
    rc, M, tr = oracle.call_model(21, params, pl, x, alm=alm, trace=True)
    assert rc == 0
    cap = int(pl[2:6].sum()) + 2
    row, nm = pkg.expand_ajAlm(params, pl, cap)
    assert nm == int(pl[2:6].sum())
    rc, M2, tr2 = oracle.mode_table_model(row, int(pl[8]), 0, x, trace=True)
    assert rc == 0
    for a, b in zip(tr, tr2):
        assert np.array_equal(a, b)                       # same windows in the same call order
    assert np.max(np.abs(M2 - M) / np.abs(M)) < 1e-13


def test_host_expander_rejects_what_the_reference_exits_on(pkg):
    rng = np.random.default_rng(1)
    params, pl = pkg.synth.ajalm_params(rng, Nmax=4, lmax=2, filter_code=1)        # "gauss": models.cpp:1451-1457
    with pytest.raises(pkg.TamcmcError) as e:
        pkg.expand_ajAlm(params, pl, 20)
    assert e.value.status == pkg.ERR_MODEL
    params, pl = pkg.synth.ajalm_params(rng, Nmax=4, lmax=2, decompose_Alm=3)      # models.cpp:1601-1603
    with pytest.raises(pkg.TamcmcError):
        pkg.expand_ajAlm(params, pl, 20)
    params, pl = pkg.synth.ajalm_params(rng, Nmax=4, lmax=2)
    with pytest.raises(pkg.TamcmcError):
        pkg.expand_ajAlm(params, pl, 5)                                             # capacity too small


@pytest.mark.gpu
@pytest.mark.parametrize("decompose", [-1, 1])
def test_gpu_ajalm_matches_oracle(pkg, oracle, decompose):
    """C3-shaped case at a size the oracle finishes in seconds: kplr-like 33-mode l<=2 list, gate filter."""
    rng = np.random.default_rng(3)
    params, pl = pkg.synth.ajalm_params(rng, Nmax=11, lmax=2, decompose_Alm=decompose, asym=0.0 if decompose == 1 else 15.0)
    x = pkg.synth.freq_axis(120000, 1900.0)
    alm = lambda l, m, t0, de, fc, user: pkg.host_alm(l, m, t0, de, fc)
    rc, M, tr = oracle.call_model(21, params, pl, x, alm=alm, trace=True)
    assert rc == 0
    y = pkg.synth.chi2_2dof_spectrum(rng, M)
    Nch = 4
    P = pkg.synth.perturb_chains(rng, params, pl, Nch)
    T = pkg.synth.tcoefs(Nch, 1.7)
    L_ref = np.zeros(Nch)
    for c in range(Nch):
        rc, Mc = oracle.call_model(21, P[c], pl, x, alm=alm)
        assert rc == 0
        L_ref[c] = oracle.call_likelihood(y, Mc, 1.0, T[c])
    cap = int(pl[2:6].sum())
    rows = np.stack([pkg.expand_ajAlm(P[c], pl, cap)[0] for c in range(Nch)])
    mpl = pkg.synth.mode_table_plength(cap, int(pl[8]), 0)
    with pkg.Context(pkg.Star(pkg.synth.MODEL_MODE_TABLE, mpl, rows.shape[1], x, y), Nch, T) as ctx:
        rcw, wl, w0, w1 = ctx.windows(rows[0])
        assert np.array_equal(wl, tr[0]) and np.array_equal(w0, tr[1]) and np.array_equal(w1, tr[2])
        Mg = ctx.model(rows[0])
        assert np.max(np.abs(Mg - M) / np.abs(M)) < RTOL
        L, st = ctx.eval(rows)
        assert (st == 0).all()
        assert np.max(np.abs(L[0] - L_ref) / np.abs(L_ref)) < RTOL
