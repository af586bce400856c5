"""Pins the plain-C oracle against the REFERENCE's own sources.

Two legs:
  * live: oracle/_ref/libtamcmc_refshim.so = the reference's likelihoods.cpp, noise_models.cpp, build_lorentzian.cpp,
    function_rot.cpp, acoefs.cpp, interpol.cpp, linfit.cpp, models.cpp (+ARMM) compiled where they lie against the
    Eigen-API shim (oracle/eigen_shim).  Skipped where that library was not built.
  * fixtures: tests/golden/reference_cpp_vectors.npz, written by tests/golden/make_golden_from_reference_cpp.py from the
    same library, committed so the pin also holds where the reference tree does not exist.
Windows are compared bit-exactly; models to 1e-13 (same scalar arithmetic; only summation order inside
`.sum()` and long-double double-rounding can differ)."""
import os

import numpy as np
import pytest

import _cases
import _refshim

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "reference_cpp_vectors.npz")
needs_ref = pytest.mark.skipif(not _refshim.available(), reason="reference sources were not compiled here (oracle/_ref/libtamcmc_refshim.so)")

MODEL_TOL = 1e-13


@needs_ref
def test_scalars_match_reference(oracle):
    R = _refshim.get()
    for s in range(0, 7):
        for l in range(0, 4):
            for m in range(-l, l + 1):
                assert oracle.Pslm(s, l, m) == R.Pslm(s, l, m), (s, l, m)     # bit-exact (incl. long double -> double)
    for l in range(1, 4):
        for m in range(-l, l + 1):
            assert oracle.Qlm(l, m) == R.Qlm(l, m)
        for inc in (0.0, 3.3, 27.0, 45.0, 63.4, 89.9, 90.0):
            assert np.array_equal(oracle.amplitude_ratio(l, inc), R.amplitude_ratio(l, inc))
    rng = np.random.default_rng(0)
    fl0 = np.sort(rng.uniform(800, 3000, 15))
    W = rng.uniform(0.5, 6, 15)
    for xi in list(rng.uniform(700, 3100, 30)) + [fl0[0], fl0[-1], fl0[3]]:
        assert oracle.lin_interpol(fl0, W, xi) == R.lin_interpol(fl0, W, xi)
    # linfit uses Eigen's .sum(): summation order may differ -> 1e-14
    assert oracle.eta0_fct(fl0) == pytest.approx(R.eta0_fct(fl0), rel=1e-13)


@needs_ref
def test_windows_bit_exact_vs_reference(oracle, pkg):
    R = _refshim.get()
    rng = np.random.default_rng(1)
    x = pkg.synth.freq_axis(60000, 50.0, pkg.synth.RESOL_4YR * 3)
    step = x[1] - x[0]
    n = 0
    for _ in range(4000):
        l = int(rng.integers(0, 4))
        fc = rng.uniform(x[0] - 100, x[-1] + 100)
        g = rng.choice([rng.uniform(0.01, 1.0), 1.0, rng.uniform(1.0, 12.0)])
        fs = rng.choice([rng.uniform(-1, 1.0), 1.0, rng.uniform(1.0, 4.0)])
        c = rng.choice([5.0, 20.0, 30.0, 50.0])
        rc, a, b = oracle.set_imin_imax(x, l, fc, g, fs, c, step)
        if rc != 0:
            continue            # the reference would exit(): not callable through the shim
        assert (a, b) == R.set_imin_imax(x, l, fc, g, fs, c, step)
        n += 1
    assert n > 3000


@needs_ref
@pytest.mark.parametrize("asym", [0.0, -43.0])
def test_profiles_noise_likelihood_vs_reference(oracle, asym):
    import ctypes as C
    dp = C.POINTER(C.c_double)
    R = _refshim.get()
    rng = np.random.default_rng(2)
    x = np.linspace(1980.0, 2020.0, 4001)
    for l in range(4):
        V = oracle.amplitude_ratio(l, 61.0) if l else np.ones(1)
        H, fc, fs, eta0, a3, g = 9.1, 2000.123, 0.93, 2.7e7, -0.021, 1.37
        res = np.zeros_like(x)
        oracle.L.orc_build_l_mode_a1etaa3(x.ctypes.data_as(dp), len(x), H, fc, fs, eta0, a3, asym, g, l, V.ctypes.data_as(dp), res.ctypes.data_as(dp))
        assert np.array_equal(res, R.build_l_mode_a1etaa3(x, H, fc, fs, eta0, a3, asym, g, l, V))
        a = [0.93, 0.05, -0.02, 0.004, 0.003, -0.001]
        res2 = np.zeros_like(x)
        oracle.L.orc_build_l_mode_aj(x.ctypes.data_as(dp), len(x), H, fc, *a, eta0, asym, g, l, V.ctypes.data_as(dp), res2.ctypes.data_as(dp))
        assert np.array_equal(res2, R.build_l_mode_aj(x, H, fc, a, eta0, asym, g, l, V))
    # harvey_like + chi22p
    noise = np.array([2.0, 300.0, 1.7, 1.0, 100.0, 2.3, 0.0, 5.0, 2.0, 0.1])
    y0 = rng.uniform(0, 3, len(x))
    import ctypes
    yo = y0.copy()
    ref = R.harvey_like(noise, x, y0, 3)
    M = oracle.call_model  # noqa: F841  (oracle's harvey_like is exercised through the models below)
    model = ref
    yobs = model * rng.exponential(1.0, len(x))
    assert oracle.chi22p(yobs, model, 1) == pytest.approx(R.chi22p(yobs, model, 1), rel=1e-14)


@needs_ref
@pytest.mark.parametrize("model_id", _cases.ALL_MODELS)
def test_models_vs_reference(oracle, pkg, model_id):
    R = _refshim.get()
    for seed in range(4):
        params, pl, x = _cases.ms_case(pkg.synth, model_id, seed=seed, N=9000 + 13 * seed, asym=(0.0 if seed % 2 == 0 else 17.0),
                                       do_amp=(seed >> 1) & 1)
        rc, M = oracle.call_model(model_id, params, pl, x)
        assert rc == 0
        rcr, Mr = R.call_model(model_id, params, pl, x)
        assert rcr == 0
        assert np.max(np.abs(M - Mr) / np.abs(Mr)) < MODEL_TOL


@needs_ref
def test_chi_square_likelihood_vs_reference(oracle):
    """likelihood_chi_square (likelihoods.cpp:31-40) -- the second likelihood of call_likelihood (model_def.cpp:402-406)."""
    import ctypes as C
    dp = C.POINTER(C.c_double)
    R = _refshim.get()
    rng = np.random.default_rng(11)
    for n in (1, 2, 777, 20000):
        model = rng.uniform(0.1, 50.0, n)
        y = model + rng.normal(0, 1, n) * rng.uniform(0.5, 2.0, n)
        sigma = rng.uniform(0.3, 3.0, n)
        mine = float(oracle.L.orc_likelihood_chi_square(y.ctypes.data_as(dp), model.ctypes.data_as(dp), sigma.ctypes.data_as(dp), n))
        assert mine == pytest.approx(R.chi_square(y, model, sigma), rel=1e-14)
        ones = np.ones(n)                      # a data file without a sigma column (config.cpp:367-374)
        mine1 = float(oracle.L.orc_likelihood_chi_square(y.ctypes.data_as(dp), model.ctypes.data_as(dp), ones.ctypes.data_as(dp), n))
        assert mine1 == pytest.approx(R.chi_square(y, model, ones), rel=1e-14)


REF_GRIDS = "/root/reference/external/Alm/data/Alm_grids_CPP/1deg_grids"
needs_alm = pytest.mark.skipif(not (_refshim.available() and hasattr(_refshim.get().L, "ref_Alm") and os.path.isdir(REF_GRIDS)),
                               reason="reference Alm sources / shipped grids not available here")


@needs_alm
@pytest.mark.parametrize("filter_code", [0, 2])
@pytest.mark.parametrize("decompose", [-1, 0, 1, 2])
def test_ajalm_model_vs_reference(oracle, pkg, decompose, filter_code):
    """Oracle model 21 against the reference's own model_MS_Global_ajAlm_HarveyLike (models.cpp:1411-1746) with the grid set-up
    of Config::Config (config.cpp:77-147) on the shipped 1-degree grids.  The oracle is fed the Alm values of the reference's
    own interpolation chain (so the pin is on the model function: unpacking, decompose_Alm paths, eval_acoefs, windows)."""
    R = _refshim.get()
    assert R.alm_grids_load(REF_GRIDS) == 0
    alm = lambda l, m, t0, de, fc, user: R.Alm_interp(l, m, t0, de, fc)
    for seed in range(3):
        rng = np.random.default_rng(100 * seed + 10 * filter_code + decompose + 1)
        params, pl = pkg.synth.ajalm_params(rng, Nmax=7, lmax=(3 if seed != 1 else 2), f0=1200.0, dnu=75.0, decompose_Alm=decompose,
                                            filter_code=filter_code, asym=(0.0 if seed == 0 else 21.0), do_amp=seed & 1, trunc_c=25.0,
                                            theta0=rng.uniform(5, 85), delta=rng.uniform(1, 44), eta_switch=float(seed != 2),
                                            epsilon=rng.uniform(1e-4, 8e-3))
        x = pkg.synth.freq_axis(16000 + 11 * seed, 1100.0, 0.05)
        rc, M, tr = oracle.call_model(21, params, pl, x, alm=alm, trace=True)
        assert rc == 0
        rcr, Mr = R.call_model(21, params, pl, x)
        assert rcr == 0
        assert np.max(np.abs(M - Mr) / np.abs(Mr)) < MODEL_TOL
        # and the product's host expander, fed the product's own grid interpolation, lands on the reference's spectrum
        G = pkg.AlmGrids(REF_GRIDS)
        row, nm = pkg.expand_ajAlm(params, pl, int(pl[2:6].sum()), alm=G)
        rc, M2 = oracle.mode_table_model(row, int(pl[8]), 0, x)
        assert rc == 0
        assert np.max(np.abs(M2 - Mr) / np.abs(Mr)) < 1e-12


def test_golden_fixtures_from_reference_cpp(oracle):
    """Holds everywhere: fixtures generated from the reference-compiled library in the build container."""
    g = np.load(GOLD, allow_pickle=False)
    n = int(g["ncases"])
    assert n >= 6
    for k in range(n):
        mid = int(g["model_id_%d" % k])
        rc, M = oracle.call_model(mid, g["params_%d" % k], g["plength_%d" % k], g["x_%d" % k])
        assert rc == 0
        assert np.max(np.abs(M - g["model_%d" % k]) / np.abs(g["model_%d" % k])) < MODEL_TOL
    # windows (bit-exact)
    wx = g["win_x"]
    step = wx[1] - wx[0]
    for row, exp in zip(g["win_in"], g["win_out"]):
        l, fc, gam, fs, c = row
        rc, a, b = oracle.set_imin_imax(wx, int(l), fc, gam, fs, c, step)
        assert rc == 0 and (a, b) == (int(exp[0]), int(exp[1]))


def test_c2_fullsize_golden_from_reference(oracle, pkg):
    """BASELINE config C2 at full size (250k bins, 80 modes, 10 chains): the oracle against the log-likelihoods the REFERENCE's
    own functions gave for the same seeded inputs (tests/golden/make_golden_c2_fullsize.py)."""
    import json
    sys_path_golden = os.path.join(HERE, "golden")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_c2_fullsize", os.path.join(sys_path_golden, "make_golden_c2_fullsize.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    gold = json.load(open(os.path.join(sys_path_golden, "reference_c2_fullsize.json")))
    params, pl, x, y, P, T = mod.c2_inputs(pkg.synth, oracle, 0.0)
    gd = gold["asym_0"]
    assert y.sum() == pytest.approx(gd["y_sum"], rel=1e-13)          # the regenerated inputs are the ones the reference saw
    rc, L = oracle.eval_chains(3, P[:3], pl, x, y, T[:3])
    assert rc == 0
    assert np.max(np.abs(L - np.array(gd["logL_reference"][:3])) / np.abs(L)) < 1e-12
