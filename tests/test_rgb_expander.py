"""Red-giant host expander (the host half of BASELINE configs C1 / C4): tamcmc_host_expand_rgb_v4 and
the ARMM mixed-mode solver / bias spline behind it (tamcmc-c_b200/csrc/host_rgb.cpp), against

  * the reference's OWN functions compiled where they lie (oracle/_ref): solve_mm_asymptotic_O2p / _O2from_l0
    (external/ARMM/solver_mm.cpp), tk::spline (external/spline), and model_RGB_asympt_aj_{AppWidth,CteWidth}_HarveyLike_v4
    with every optimum_lorentzian_calc_aj call they make recorded (oracle/ref_shim_api.cpp) -- BIT FOR BIT, because a mixed
    mode narrower than 1e-3 microHz needs its frequency to ~1e-15 relative for 1e-10 on the spectrum;
  * the committed recording of the same calls on the reference's fixture 10722175 (tests/golden/reference_rgb_vectors.npz,
    written with ONE OpenMP thread: the reference's zeta sums run under `omp critical`), which holds where /root/reference
    does not exist;
  * on the GPU: reference parameter vector -> host expander -> mode table -> model spectrum and logL against the spectrum the
    reference's model function returned (1e-10)."""
import os

import numpy as np
import pytest

import _refshim

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "reference_rgb_vectors.npz")
needs_ref = pytest.mark.skipif(not (_refshim.available() and hasattr(_refshim.get().L, "ref_solve_mm_from_l0")),
                               reason="reference ARMM sources not compiled here")
COLS = ["l", "fc", "H", "W", "a1", "a2", "a3", "a4", "a5", "a6", "eta0"]


def _records(pkg, row, pl, nm):
    nn = int(pl[8])
    rec = row[4 + nn:4 + nn + 20 * nm].reshape(nm, 20)
    return rec[np.lexsort((rec[:, 1], rec[:, 0]))]


@pytest.mark.parametrize("model_id", [25, 27])
def test_rows_match_the_reference_recording_bit_for_bit(pkg, model_id):
    G = np.load(GOLD)
    x = G["x"]
    step = x[2] - x[1]
    for i in range(int(G["ncases"])):
        params, pl = G["params%d" % i], G["plength%d" % i]
        rows = G[("rows%d" if model_id == 25 else "rows27_%d") % i]
        row, nm = pkg.expand_rgb_v4(model_id, params, pl, step, 100)
        assert nm == len(rows)
        rec = _records(pkg, row, pl, nm)
        for c, name in enumerate(COLS):
            assert np.array_equal(rec[:, c], rows[:, c]), (model_id, i, name, np.max(np.abs(rec[:, c] - rows[:, c])))
        assert np.all(rec[:, 11:] == 0.0)                                           # no extra shifts, padding untouched
        # header: nmodes, inclination, trunc_c, asym; then the raw noise parameters (abs() is applied on the device, models.cpp:5017)
        o = int(pl[:8].sum())
        assert row[0] == nm and row[1] == abs(params[o + int(pl[8])]) and row[2] == rows[0, 13] and row[3] == rows[0, 11]
        assert np.array_equal(row[4:4 + int(pl[8])], params[o:o + int(pl[8])])
        # the reference's own call order: all l=0, the mixed l=1 modes by increasing frequency, l=2, l=3 (models.cpp:4931-5006)
        nn = int(pl[8])
        ls = row[4 + nn:4 + nn + 20 * nm].reshape(nm, 20)[:, 0]
        assert np.array_equal(ls, np.sort(ls))
    assert pkg.expand_rgb_v4(model_id, params, pl, step, 100)[1] == nm          # deterministic whatever the thread schedule
    assert np.array_equal(pkg.expand_rgb_v4(model_id, params, pl, step, 100)[0], row)


def test_expander_errors(pkg):
    G = np.load(GOLD)
    x = G["x"]
    step = x[2] - x[1]
    params, pl = G["params0"], G["plength0"]
    with pytest.raises(pkg.TamcmcError) as e:
        pkg.expand_rgb_v4(25, params, pl, step, 20)                                  # capacity too small
    assert e.value.status == pkg.ERR_ARG
    with pytest.raises(pkg.TamcmcError) as e:
        pkg.expand_rgb_v4(24, params, pl, step, 100)                                 # obsolete id: the reference exits (model_def.cpp:336-340)
    assert e.value.status == pkg.ERR_MODEL
    p2 = params.copy()
    p2[8:13] = [3.0, 12.0, 21.0, 30.0, 39.0]                                         # fmin - Dnu < 0: "THE ARMM WILL NOT CONVERGE" (models.cpp:4852-4858)
    with pytest.raises(pkg.TamcmcError) as e:
        pkg.expand_rgb_v4(25, p2, pl, step, 100)
    assert e.value.status == pkg.ERR_NONFINITE
    p3 = params.copy()
    p3[-2] = 1.0                                                                     # cubic-spline bias with all-zero fref differences is fine; Nferr < 3 is not
    p3[-1] = 2.0
    pl3 = pl.copy()
    with pytest.raises(pkg.TamcmcError):
        pkg.expand_rgb_v4(25, p3, pl3, step, 100)


@needs_ref
def test_armm_solver_and_spline_against_the_reference_functions(pkg):
    R = _refshim.get()
    R.single_thread()
    rng = np.random.default_rng(21)
    resol = 1e6 / (4 * 365. * 86400.)
    for k in range(6):
        Dnu = rng.uniform(4.0, 14.0)
        n = int(rng.integers(4, 8))
        f0 = rng.uniform(60.0, 180.0)
        fl0 = f0 + Dnu * np.arange(n) + rng.uniform(-0.02, 0.02, n) * Dnu
        d01, DP, alpha, q = rng.uniform(-0.3, 0.3), rng.uniform(60.0, 90.0), rng.uniform(0.0, 0.9), rng.uniform(0.05, 0.3)
        mine = pkg.armm_solve_from_l0(fl0, 1, d01, DP, alpha, q, resol * (1 + k % 2), fl0.min(), fl0.max())
        ref = R.solve_mm_from_l0(fl0, 1, d01, DP, alpha, q, resol * (1 + k % 2), fl0.min(), fl0.max())
        assert len(ref[0]) > 10
        for a, b in zip(mine, ref):
            assert np.array_equal(a, b)
        eps = rng.uniform(0.0, 1.0)
        mine = pkg.armm_solve_O2p(Dnu, eps, 1, d01 / Dnu, 0.0, 0.0, DP, alpha, q, fl0.min() - Dnu, fl0.max() + Dnu, resol)
        ref = R.solve_mm_O2p(Dnu, eps, 1, d01 / Dnu, 0.0, 0.0, DP, alpha, q, fl0.min() - Dnu, fl0.max() + Dnu, resol)
        assert len(ref[0]) > 10
        for a, b in zip(mine, ref):
            assert np.array_equal(a, b)
    # a sub-giant: few g modes -> the wide search zone (solver_mm.cpp:509-515)
    fl0 = 400.0 + 40.0 * np.arange(6)
    mine = pkg.armm_solve_from_l0(fl0, 1, 1.0, 400.0, 0.3, 0.2, resol * 4, fl0.min(), fl0.max())
    ref = R.solve_mm_from_l0(fl0, 1, 1.0, 400.0, 0.3, 0.2, resol * 4, fl0.min(), fl0.max())
    assert len(ref[0]) >= 3 and all(np.array_equal(a, b) for a, b in zip(mine, ref))
    # tk::spline, both kinds, inside and outside the knots
    for kind in (1, 2):
        xs = np.sort(rng.uniform(80.0, 130.0, 9))
        ys = rng.uniform(-0.05, 0.05, 9)
        xq = np.concatenate([rng.uniform(70.0, 140.0, 200), xs])
        assert np.array_equal(pkg.spline_eval(xs, ys, xq, kind), R.spline_eval(xs, ys, xq, kind))
    assert np.allclose(pkg.spline_eval([0.0, 1.0, 2.0, 3.0], [0.0, 1.0, 2.0, 3.0], [-1.0, 0.5, 4.0], 1), [-1.0, 0.5, 4.0], atol=1e-15)


@needs_ref
@pytest.mark.parametrize("model_id", [25, 27])
def test_rows_match_the_live_reference_on_perturbed_parameters(pkg, model_id):
    """Random walks around the fixture's parameters (what the chains of an MCMC visit): same mode count, same rows."""
    R = _refshim.get()
    R.single_thread()
    G = np.load(GOLD)
    x = G["x"]
    step = x[2] - x[1]
    rng = np.random.default_rng(5 + model_id)
    for i in range(int(G["ncases"])):
        params, pl = G["params%d" % i].copy(), G["plength%d" % i]
        o_l1 = int(pl[0] + pl[1] + pl[2])
        params[8:13] += rng.normal(0, 0.05, 5)                       # l=0 frequencies
        params[o_l1 + 1] *= 1 + rng.normal(0, 0.01)                  # DP1
        params[o_l1 + 3] *= 1 + rng.normal(0, 0.05)                  # q
        params[o_l1] = rng.normal(0, 0.1)                            # delta01
        params[:5] *= rng.uniform(0.8, 1.2, 5)                       # heights
        rc, M, rows = R.call_model_recorded(model_id, params, pl, x)
        assert rc == 0
        rows = rows[np.lexsort((rows[:, 1], rows[:, 0]))]
        row, nm = pkg.expand_rgb_v4(model_id, params, pl, step, 120)
        assert nm == len(rows)
        rec = _records(pkg, row, pl, nm)
        for c, name in enumerate(COLS):
            assert np.array_equal(rec[:, c], rows[:, c]), (i, name)


@pytest.mark.gpu
@pytest.mark.parametrize("model_id", [25, 27])
def test_gpu_from_reference_parameter_vector(pkg, oracle, model_id):
    """ids 25 / 27 from a REFERENCE parameter vector: host expander -> mode table -> GPU model and logL against the spectrum the
    reference's model function returned for that vector (fixture 10722175, 6099 bins)."""
    G = np.load(GOLD)
    x, y = G["x"], G["y"]
    step = x[2] - x[1]
    n = int(G["ncases"])
    rows, models = [], []
    cap = 100
    for i in range(n):
        params, pl = G["params%d" % i], G["plength%d" % i]
        rows.append(pkg.expand_rgb_v4(model_id, params, pl, step, cap)[0])
        models.append(G[("model%d" if model_id == 25 else "model27_%d") % i])
    rows = np.stack(rows)
    T = pkg.synth.tcoefs(n, 3.5)
    mpl = pkg.synth.mode_table_plength(cap, int(pl[8]), 1)
    with pkg.Context(pkg.Star(pkg.synth.MODEL_MODE_TABLE, mpl, rows.shape[1], x, y), n, T) as ctx:
        L, st = ctx.eval(rows)
        assert (st == 0).all()
        for i in range(n):
            M = ctx.model(rows[i])
            assert np.max(np.abs(M - models[i]) / np.abs(models[i])) < 1e-10
            Lr = oracle.call_likelihood(y, models[i], 1.0, T[i])
            assert abs(L[0, i] - Lr) <= 1e-10 * abs(Lr)


def test_bisection_between_the_tangent_poles_gives_the_scan_rows(pkg, tmp_path):
    """The mixed-mode solver locates the sign changes of p - g by bisection between the poles of the tangent and brackets each
    root by bisection on the local grid (host_rgb.cpp: band_sign_changes, interp_zero_lazy); TAMCMC_ARMM_EXACT_SCAN=1 evaluates
    every grid point and walks to the bracket like the reference does.  Same rows bit for bit, on the four recorded vectors and on
    perturbed ones (other DP1, q, alpha_g, delta01: other pole positions)."""
    import subprocess
    import sys
    G = np.load(GOLD)
    step = float(G["x"][2] - G["x"][1])
    rng = np.random.default_rng(77)
    cases = []
    for i in range(int(G["ncases"])):
        params, pl = G["params%d" % i], G["plength%d" % i]
        cases.append((params, pl))
        for _ in range(3):
            p = params.copy()
            o1 = int(pl[:2].sum()) + int(pl[2])             # l=1 block: delta01, DP1, alpha_g, q
            p[o1] = rng.uniform(-0.05, 0.05); p[o1 + 1] *= rng.uniform(0.93, 1.07); p[o1 + 2] = rng.uniform(0.0, 0.9); p[o1 + 3] = rng.uniform(0.05, 0.4)
            cases.append((p, pl))
    np.savez(tmp_path / "cases.npz", **{"p%d" % k: c[0] for k, c in enumerate(cases)}, **{"l%d" % k: c[1] for k, c in enumerate(cases)}, n=len(cases), step=step)
    script = ("import sys, numpy as np\nsys.path.insert(0, %r)\nimport __graft_entry__ as g\npkg = g.load_package()\n"
              "C = np.load(%r)\nout = {}\n"
              "for k in range(int(C['n'])):\n    row, nm = pkg.expand_rgb_v4(25, C['p%%d' %% k], C['l%%d' %% k], float(C['step']), 120)\n    out['r%%d' %% k] = row\n"
              "np.savez(%r, **out)\n") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(tmp_path / "cases.npz"), str(tmp_path / "exact.npz"))
    env = dict(os.environ, TAMCMC_ARMM_EXACT_SCAN="1")
    subprocess.run([sys.executable, "-c", script], check=True, env=env)
    E = np.load(tmp_path / "exact.npz")
    for k, (p, pl) in enumerate(cases):
        row, nm = pkg.expand_rgb_v4(25, p, pl, step, 120)
        assert np.array_equal(row, E["r%d" % k]), k
