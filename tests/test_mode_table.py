"""Generic mode table (TAMCMC_MODEL_MODE_TABLE): the GPU entry for model functions whose mode list is resolved on the
host.  Pinned on the red-giant mixed-mode model: tests/golden/reference_rgb_vectors.npz holds, for the reference's own
fixture 10722175 (BASELINE configs C1/C4), the spectra returned by the REFERENCE's model_RGB_asympt_aj_AppWidth_HarveyLike_v4
and the optimum_lorentzian_calc_aj calls it made (tests/golden/make_golden_rgb_from_reference_cpp.py)."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "reference_rgb_vectors.npz")
RTOL = 1e-10      # north_star: model spectrum and logL within 1e-10 relative (FP64)


def _cases(pkg):
    """-> x, y, [(row, Nnoise, reference model, rows)] with one common table capacity."""
    g = np.load(GOLD)
    x, y = g["x"], g["y"]
    n = int(g["ncases"])
    cap = max(len(g["rows%d" % i]) for i in range(n)) + 3
    out = []
    for i in range(n):
        params, pl, rows, M = g["params%d" % i], g["plength%d" % i], g["rows%d" % i], g["model%d" % i]
        o_noise = int(pl[:8].sum())
        Nnoise = int(pl[8])
        noise = params[o_noise:o_noise + Nnoise]
        inc = abs(params[o_noise + Nnoise])                      # models.cpp:4768
        # every call of one model evaluation shares asym, step and trunc_c
        assert np.all(rows[:, 11] == rows[0, 11]) and np.all(rows[:, 13] == rows[0, 13])
        assert np.all(rows[:, 12] == x[2] - x[1])                # RGB v4: step = x[2]-x[1] (models.cpp:4714)
        row = pkg.synth.mode_table_row(cap, inc, rows[0, 13], rows[0, 11], noise, rows[:, :11])
        out.append((row, Nnoise, M, rows, inc))
    return x, y, cap, out


def test_oracle_mode_table_matches_reference_rgb_model(pkg, oracle):
    x, y, cap, cases = _cases(pkg)
    for row, Nnoise, M_ref, rows, inc in cases:
        # the m-height ratios the reference used are amplitude_ratio(l, inclination) (bit-exact)
        for r in rows:
            l = int(r[0])
            V = oracle.amplitude_ratio(l, inc) if l > 0 else np.array([1.0])
            assert np.array_equal(V, r[14:14 + 2 * l + 1])
        rc, M = oracle.mode_table_model(row, Nnoise, 1, x)
        assert rc == 0
        assert np.max(np.abs(M - M_ref) / np.abs(M_ref)) < 1e-12


def test_mode_table_row_layout(pkg):
    row = pkg.synth.mode_table_row(3, 45.0, 25.0, 0.0, [1.0, 2.0, 3.0, 0.1], [[1, 100.0, 2.0, 0.5, 0.3, 0, 0, 0, 0, 0, 0]])
    assert len(row) == pkg.synth.mode_table_nparams(3, 4) == 4 + 4 + 60
    assert list(row[:8]) == [1, 45.0, 25.0, 0.0, 1.0, 2.0, 3.0, 0.1]
    assert list(row[8:13]) == [1, 100.0, 2.0, 0.5, 0.3] and not row[19:].any()


@pytest.mark.gpu
def test_gpu_mode_table_matches_reference_rgb_model(pkg, oracle):
    x, y, cap, cases = _cases(pkg)
    Nnoise = cases[0][1]
    rows = np.stack([c[0] for c in cases])
    T = pkg.synth.tcoefs(len(cases), 3.5)              # config_default.cfg: Tcoef lambda of the quick-start preset
    rc, L_ref = oracle.mode_table_eval_chains(rows, Nnoise, 1, x, y, T)
    assert rc == 0
    pl = pkg.synth.mode_table_plength(cap, Nnoise, step_mode=1)
    star = pkg.Star(pkg.synth.MODEL_MODE_TABLE, pl, rows.shape[1], x, y)
    with pkg.Context(star, len(cases), T) as ctx:
        for row, _, M_ref, _, _ in cases:
            rc, M_orc, tr = oracle.mode_table_model(row, Nnoise, 1, x, trace=True)
            rcw, wl, w0, w1 = ctx.windows(row)
            assert rcw == 0
            assert np.array_equal(wl, tr[0]) and np.array_equal(w0, tr[1]) and np.array_equal(w1, tr[2])   # bit-exact
            Mg = ctx.model(row)
            assert np.max(np.abs(Mg - M_ref) / np.abs(M_ref)) < RTOL     # against the REFERENCE's own output
            assert np.max(np.abs(Mg - M_orc) / np.abs(M_orc)) < RTOL
        L, st = ctx.eval(rows)
        assert (st == 0).all()
        assert np.max(np.abs(L[0] - L_ref) / np.abs(L_ref)) < RTOL
        L2, _ = ctx.eval(rows[::-1].copy())            # different per-chain mode counts in other slots: no stale records
        assert np.allclose(L2[0][::-1] * T[::-1], L[0] * T, rtol=1e-14, atol=0)


@pytest.mark.gpu
def test_gpu_mode_table_extra_shifts_and_errors(pkg, oracle):
    """Per-m extra frequency shifts (the fc*epsilon*Alm term of build_l_mode_ajAlm, build_lorentzian.cpp:182-190) and the
    argument / record validation of the table."""
    rng = np.random.default_rng(5)
    x = pkg.synth.freq_axis(20000, 900.0, 0.02)
    noise = [1.0, 100.0, 2.0, 0.5, 10.0, 2.0, 0.1]
    nm = 9
    modes = np.zeros((nm, 11))
    modes[:, 0] = [0, 1, 2, 3, 0, 1, 2, 3, 1]
    modes[:, 1] = np.sort(rng.uniform(950, 1250, nm))
    modes[:, 2] = rng.uniform(1, 20, nm)
    modes[:, 3] = rng.uniform(0.3, 3, nm)
    modes[:, 4] = rng.uniform(0.3, 2, nm)          # a1
    modes[:, 6] = rng.uniform(-0.05, 0.05, nm)     # a3
    modes[:, 8] = rng.uniform(-0.01, 0.01, nm)     # a5
    modes[:, 10] = 3e-7 * (rng.uniform(size=nm) > 0.5)
    extra = rng.uniform(-0.3, 0.3, (nm, 7))
    extra[modes[:, 0] == 0] = 0
    row = pkg.synth.mode_table_row(12, 51.0, 20.0, 14.0, noise, modes, extra)
    rc, M, tr = oracle.mode_table_model(row, len(noise), 0, x, trace=True)
    assert rc == 0
    pl = pkg.synth.mode_table_plength(12, len(noise), 0)
    with pkg.Context(pkg.Star(pkg.synth.MODEL_MODE_TABLE, pl, len(row), x, M), 2, [1.0, 2.0]) as ctx:
        rcw, wl, w0, w1 = ctx.windows(row)
        assert np.array_equal(wl, tr[0]) and np.array_equal(w0, tr[1]) and np.array_equal(w1, tr[2])
        Mg = ctx.model(row)
        assert np.max(np.abs(Mg - M) / np.abs(M)) < RTOL
        bad = row.copy(); bad[0] = 13                      # more modes than the capacity
        neg = row.copy(); neg[4 + len(noise) + 2] = -1.0   # negative height
        L, st = ctx.eval(np.stack([row, bad]), raise_on_error=False)
        assert st[0, 0] == 0 and (st[0, 1] & pkg.CHAIN_BADCFG) and np.isnan(L[0, 1]) and np.isfinite(L[0, 0])
        L, st = ctx.eval(np.stack([neg, row]), raise_on_error=False)
        assert (st[0, 0] & pkg.CHAIN_BADCFG) and st[0, 1] == 0
    with pytest.raises(pkg.TamcmcError):
        pl_bad = pl.copy(); pl_bad[3] = 2
        pkg.Context(pkg.Star(pkg.synth.MODEL_MODE_TABLE, pl_bad, len(row), x, M), 1, [1.0])


@pytest.mark.gpu
def test_gpu_dense_lists_stream_through_one_slot_in_segments(pkg, oracle):
    """Tile lists far larger than one ring slot (352 mask-free / 24 general entries): 220 modes (l up to 3 -> ~1100
    components) whose windows all cover a 2-tile spectrum, plus 70 narrow-window modes whose edges fall inside the tiles
    (several hundred general entries).  The producer streams such a list through its slot as many segments."""
    rng = np.random.default_rng(77)
    N = 2500
    x = pkg.synth.freq_axis(N, 1000.0, 0.01)            # 25 microHz wide: 2 tiles
    noise = [0.5, 200.0, 2.0, 0.2]
    nwide, nedge = 220, 70
    modes = np.zeros((nwide + nedge, 11))
    modes[:, 0] = rng.integers(0, 4, nwide + nedge)
    modes[:nwide, 1] = rng.uniform(1001, 1024, nwide)
    modes[:nwide, 3] = rng.uniform(0.2, 1.5, nwide)      # trunc_c * (l + Gamma) >> 25 microHz: every window covers both tiles
    modes[nwide:, 1] = rng.uniform(1002, 1023, nedge)
    modes[nwide:, 3] = rng.uniform(0.01, 0.05, nedge)
    modes[:, 2] = rng.uniform(0.5, 30, nwide + nedge)
    modes[:, 4] = rng.uniform(0.05, 0.4, nwide + nedge)  # a1 < 1: window half-width = c*(l+1) for narrow modes
    cap = nwide + nedge
    for asym, trunc_wide in ((0.0, 40.0), (25.0, 40.0)):
        # two chains: one with wide windows for everything, one where the narrow modes get c = 1.5 (edges inside the tiles)
        rows = np.stack([pkg.synth.mode_table_row(cap, 37.0, trunc_wide, asym, noise, modes),
                         pkg.synth.mode_table_row(cap, 37.0, 1.5, asym, noise, modes)])
        T = [1.0, 1.7]
        refs = []
        for r in rows:
            rc, M, tr = oracle.mode_table_model(r, len(noise), 0, x, trace=True)
            assert rc == 0
            refs.append((M, tr))
        w0 = refs[1][1][1]; w1 = refs[1][1][2]
        assert np.sum((w0 > 0) & (w0 < N)) + np.sum((w1 > 0) & (w1 < N)) > 100      # many window edges inside the spectrum
        rng2 = np.random.default_rng(5)
        y = pkg.synth.chi2_2dof_spectrum(rng2, refs[0][0])
        rc, L_ref = oracle.mode_table_eval_chains(rows, len(noise), 0, x, y, T)
        pl = pkg.synth.mode_table_plength(cap, len(noise), 0)
        with pkg.Context(pkg.Star(pkg.synth.MODEL_MODE_TABLE, pl, rows.shape[1], x, y), 2, T) as ctx:
            for r, (M, tr) in zip(rows, refs):
                rcw, wl, a, b = ctx.windows(r)
                assert np.array_equal(wl, tr[0]) and np.array_equal(a, tr[1]) and np.array_equal(b, tr[2])
                Mg = ctx.model(r)
                assert np.max(np.abs(Mg - M) / np.abs(M)) < RTOL
            L, st = ctx.eval(rows)
            assert (st == 0).all()
            assert np.max(np.abs(L[0] - L_ref) / np.abs(L_ref)) < RTOL
            L2, _ = ctx.eval(rows)
            assert np.array_equal(L, L2)
