"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs.  Bars: bin windows and mode indices BIT-EXACT; model spectrum and log-likelihood within
1e-10 relative (FP64 path, BASELINE.json north_star).  Run with `-m gpu` on a B200."""
import numpy as np
import pytest

import _cases

pytestmark = pytest.mark.gpu

RTOL = 1e-10  # north_star: "model spectrum and log-likelihood within 1e-10 relative in FP64 mode"


def _ctx(pkg, model_id, params, pl, x, y, Nchains=1, T=None):
    T = np.ones(Nchains) if T is None else T
    return pkg.Context(pkg.Star(model_id, pl, len(params), x, y), Nchains, T)


@pytest.mark.parametrize("model_id", _cases.ALL_MODELS)
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_windows_model_logl(pkg, oracle, model_id, seed):
    asym = 0.0 if seed % 2 == 0 else (31.0 if seed == 1 else -55.0)
    params, pl, x = _cases.ms_case(pkg.synth, model_id, seed=seed, N=30000 + 517 * seed, asym=asym, do_amp=(seed >> 1) & 1)
    rc, M, tr = oracle.call_model(model_id, params, pl, x, trace=True)
    assert rc == 0
    rng = np.random.default_rng(100 + seed)
    y = pkg.synth.chi2_2dof_spectrum(rng, M)
    P = pkg.synth.perturb_chains(rng, params, pl, 4)
    T = pkg.synth.tcoefs(4, 1.7)
    rc, L_ref = oracle.eval_chains(model_id, P, pl, x, y, T)
    assert rc == 0
    with _ctx(pkg, model_id, params, pl, x, y, 4, T) as ctx:
        rcw, wl, w0, w1 = ctx.windows(params)
        assert rcw == 0
        assert np.array_equal(wl, tr[0]) and np.array_equal(w0, tr[1]) and np.array_equal(w1, tr[2])
        Mg = ctx.model(params)
        assert np.max(np.abs(Mg - M) / np.abs(M)) < RTOL
        L, st = ctx.eval(P)
        assert (st == 0).all()
        assert np.max(np.abs(L[0] - L_ref) / np.abs(L_ref)) < RTOL
        # run-to-run bitwise reproducibility (deterministic reduction order)
        L2, _ = ctx.eval(P)
        assert np.array_equal(L, L2)


def test_component_table_matches_oracle_scalars(pkg, oracle):
    params, pl, x = _cases.ms_case(pkg.synth, 23, seed=5, N=20000)
    with _ctx(pkg, 23, params, pl, x, np.ones_like(x)) as ctx:
        mi, m, nu, h, w = ctx.components(params)
    Nmax, lmax = int(pl[0]), int(pl[1])
    o = Nmax + lmax
    fl0 = params[o:o + Nmax]
    Nf = Nmax * (lmax + 1)
    at = params[o + Nf:o + Nf + 12]
    eta0 = oracle.eta0_fct(fl0) if params[o + Nf + 12] == 1 else 0.0
    # l-major order: mode index j -> (l, n)
    for k in range(len(nu)):
        l, n = divmod(int(mi[k]), Nmax)
        fc = params[o + l * Nmax + n]
        if l == 0:
            assert nu[k] == fc
            continue
        a = [at[2 * j] + at[2 * j + 1] * (fc * 1e-3) if j < 2 * l else 0.0 for j in range(6)]
        ref = fc + sum(a[j] * oracle.Pslm(j + 1, l, int(m[k])) for j in range(6))
        if eta0 > 0:
            ref += fc * eta0 * oracle.Qlm(l, int(m[k])) * (a[0] * 1e-6) ** 2
        assert abs(nu[k] - ref) <= 2 * np.spacing(ref)


def test_edge_windows_bit_exact(pkg, oracle):
    """Modes hanging over both ends of the spectrum, widths and splittings straddling the 1.0
    thresholds of set_imin_imax (build_lorentzian.cpp:599-643)."""
    rng = np.random.default_rng(42)
    for trial in range(6):
        N = 9000 + 1000 * trial
        x = pkg.synth.freq_axis(N, 300.0, 0.05)
        params, pl = pkg.synth.classic_params(rng, Nmax=5, lmax=3, f0=260.0 + 10 * trial, dnu=(x[-1] - x[0]) / 3.5,
                                              a1=[0.3, 1.0, 1.7, 0.99999, 1.00001, 2.2][trial], trunc_c=[5, 10, 30, 50, 2, 100][trial],
                                              wmin=0.2, wmax=2.5)
        o_w = 5 + 3 + 20 + 6
        params[o_w] = 1.0                      # gamma == 1 exactly: two branches fire
        rc, M, tr = oracle.call_model(3, params, pl, x, trace=True)
        with _ctx(pkg, 3, params, pl, x, np.ones_like(x)) as ctx:
            rcw, wl, w0, w1 = ctx.windows(params, raise_on_error=False)
            assert (rcw == 0) == (rc == 0)
            assert np.array_equal(wl[: len(tr[0])], tr[0]) and np.array_equal(w0[: len(tr[0])], tr[1]) and np.array_equal(w1[: len(tr[0])], tr[2])
            if rc == 0:
                Mg = ctx.model(params)
                assert np.max(np.abs(Mg - M) / np.abs(M)) < RTOL


def test_reference_unit_test_recipe(pkg, oracle):
    """The reference's own unit-test input recipe (test_build_l_mode.cpp:100-170: make_params_aj_model,
    x = LinSpaced(0, Nfreqs*130+250) at the 4-year Kepler resolution, trunc_c = 50), seeded."""
    rng = np.random.default_rng(2024)
    resol = pkg.synth.RESOL_4YR
    for i in range(4):
        lmax = int(rng.integers(2, 4))
        Dnu = rng.uniform(129, 130.0)
        eps = rng.uniform(0, 0.05)
        d0l = rng.uniform(-Dnu * 0.02, 0)
        params, pl = pkg.synth.make_params_aj_model(rng, lmax, 5, Dnu, eps, d0l, asym_on=bool(i % 2))
        fmax = 5 * 130 + 250
        Ndata = int(np.ceil(fmax / resol))
        x = np.linspace(0.0, fmax, Ndata)
        rc, M, tr = oracle.call_model(23, params, pl, x, trace=True)
        assert rc == 0
        with _ctx(pkg, 23, params, pl, x, np.ones_like(x)) as ctx:
            rcw, wl, w0, w1 = ctx.windows(params)
            assert np.array_equal(wl, tr[0]) and np.array_equal(w0, tr[1]) and np.array_equal(w1, tr[2])
            Mg = ctx.model(params)
        # the reference's own acceptance threshold is norm(new-ref) <= 1e-8 (test_build_l_mode.cpp:104,134)
        assert np.linalg.norm(Mg - M) <= 1e-8
        assert np.max(np.abs(Mg - M) / np.abs(M)) < RTOL


def test_inactive_mask_and_status(pkg, oracle):
    params, pl, x = _cases.ms_case(pkg.synth, 3, seed=9, N=12000)
    rc, M = oracle.call_model(3, params, pl, x)
    y = M.copy()
    P = np.tile(params, (5, 1))
    T = pkg.synth.tcoefs(5, 2.0)
    rc, L_ref = oracle.eval_chains(3, P, pl, x, y, T)
    with _ctx(pkg, 3, params, pl, x, y, 5, T) as ctx:
        L, st = ctx.eval(P, active=[1, 0, 1, 0, 1])
        assert list(st[0]) == [0, pkg.CHAIN_INACTIVE, 0, pkg.CHAIN_INACTIVE, 0]
        assert np.isnan(L[0, 1]) and np.isnan(L[0, 3])
        assert np.allclose(L[0, [0, 2, 4]], L_ref[[0, 2, 4]], rtol=RTOL, atol=0)
        # tempering: same parameters, logL scales as 1/T (model_def.cpp:401)
        L_all, _ = ctx.eval(P)
        assert np.allclose(L_all[0] * T, L_all[0, 0], rtol=1e-14, atol=0)


def test_window_error_is_reported_not_fatal(pkg, oracle):
    """A negative trunc_c makes imax-imin<=0: the reference prints and exits (build_lorentzian.cpp:650-665);
    the ABI returns TAMCMC_ERR_WINDOW, flags the chain and gives NaN, other chains are evaluated."""
    params, pl, x = _cases.ms_case(pkg.synth, 3, seed=3, N=8000)
    rc, M = oracle.call_model(3, params, pl, x)
    assert rc == 0
    bad = params.copy()
    bad[-2] = -5.0
    rc_bad, _ = oracle.call_model(3, bad, pl, x)
    assert rc_bad != 0
    P = np.stack([params, bad, params])
    with _ctx(pkg, 3, params, pl, x, M, 3) as ctx:
        L, st = ctx.eval(P, raise_on_error=False)
        assert ctx.last_rc == pkg.ERR_WINDOW
        assert st[0, 0] == 0 and st[0, 2] == 0 and (st[0, 1] & pkg.CHAIN_WINDOW)
        assert np.isnan(L[0, 1]) and np.isfinite(L[0, 0]) and L[0, 0] == L[0, 2]
        # NaN parameters are data, not errors to die on (MALA.cpp:490,522): NaN logL, flagged
        nanp = params.copy()
        nanp[int(pl[0]) + int(pl[1]) + 1] = np.nan
        L, st = ctx.eval(np.stack([params, nanp, params]), raise_on_error=False)
        assert np.isnan(L[0, 1]) and np.isfinite(L[0, 0])


def test_unknown_model_and_bad_sizes(pkg):
    x = pkg.synth.freq_axis(2048, 900.0, 0.1)
    params, pl = pkg.synth.classic_params(np.random.default_rng(0), Nmax=4, lmax=2, f0=950.0, dnu=40.0)
    for mid in (2, 4, 5, 18, 19, 99):     # obsolete / unknown / self-terminating ids exit in the reference (model_def.cpp:231-384, models.cpp:599-603)
        with pytest.raises(pkg.TamcmcError) as ei:
            pkg.Context(pkg.Star(mid, pl, len(params), x, np.ones_like(x)), 1, [1.0])
        assert ei.value.status == pkg.ERR_MODEL
    with pytest.raises(pkg.TamcmcError) as ei:
        pkg.Context(pkg.Star(3, pl, len(params) - 5, x, np.ones_like(x)), 1, [1.0])
    assert ei.value.status == pkg.ERR_ARG
    with pytest.raises(pkg.TamcmcError) as ei:
        pkg.Context(pkg.Star(3, pl, len(params), x, np.ones_like(x)), 1, [1.0], likelihood_id=2)
    assert ei.value.status == pkg.ERR_LIKELIHOOD


def test_batch_of_stars_matches_single(pkg, oracle):
    """Ragged batch: stars with different N, mode counts and models in ONE launch."""
    stars, Ps, refs = [], [], []
    T = pkg.synth.tcoefs(3, 1.7)
    for s, (mid, N) in enumerate([(3, 5000), (23, 12345), (3, 1024), (12, 3073), (23, 2049)]):
        params, pl, x = _cases.ms_case(pkg.synth, mid, seed=20 + s, N=N, Nmax=3 + s, lmax=2 + (s % 2))
        rc, M = oracle.call_model(mid, params, pl, x)
        assert rc == 0
        rng = np.random.default_rng(s)
        y = pkg.synth.chi2_2dof_spectrum(rng, M)
        P = pkg.synth.perturb_chains(rng, params, pl, 3)
        rc, Lr = oracle.eval_chains(mid, P, pl, x, y, T)
        assert rc == 0
        stars.append(pkg.Star(mid, pl, len(params), x, y))
        Ps.append(P)
        refs.append(Lr)
    with pkg.Context(stars, 3, T) as ctx:
        L, st = ctx.eval(Ps)
        assert (st == 0).all()
        assert np.max(np.abs(L - np.array(refs)) / np.abs(np.array(refs))) < RTOL
    for s in (1, 3):
        with pkg.Context(stars[s], 3, T) as ctx1:
            L1, _ = ctx1.eval(Ps[s])
            assert np.array_equal(L1[0], L[s])     # batching does not change a star's result bitwise


def test_bin_sharded_sums_add_up(pkg, oracle):
    """A spectrum split by bins over 'ranks': per-shard raw sums S add up to the whole."""
    import torch
    params, pl, x = _cases.ms_case(pkg.synth, 3, seed=11, N=50000)
    rc, M = oracle.call_model(3, params, pl, x)
    rng = np.random.default_rng(3)
    y = pkg.synth.chi2_2dof_spectrum(rng, M)
    P = pkg.synth.perturb_chains(rng, params, pl, 4)
    T = pkg.synth.tcoefs(4, 1.7)
    rc, L_ref = oracle.eval_chains(3, P, pl, x, y, T)
    from importlib import import_module
    shard = import_module("tamcmc_c_b200.sharding")
    tr = oracle.call_model(3, params, pl, x, trace=True)[2]
    work = shard.bin_work(len(x), *tr)
    for world in (2, 3):
        ranges = shard.bin_shards(len(x), world, work)
        assert ranges[0][0] == 0 and ranges[-1][1] == len(x)
        S = torch.zeros(4, dtype=torch.float64, device="cuda")
        for lo, hi in ranges:
            assert lo % shard.TILE == 0
            with pkg.Context(pkg.Star.shard(3, pl, len(params), x, y, lo, hi), 4, T) as ctx:
                dP = torch.tensor(ctx.pack_params(P), device="cuda")
                dS = torch.zeros(4, dtype=torch.float64, device="cuda")
                ctx.eval_device(dP.data_ptr(), dS.data_ptr(), raw_sum=True)
                ctx.sync()
                S += dS
        L = shard.finalize_logL(S.cpu().numpy(), 1.0, T)
        assert np.max(np.abs(L - L_ref) / np.abs(L_ref)) < RTOL


def test_extreme_dynamic_range_components(pkg, oracle):
    """Heights far outside the fast-path range and very narrow modes go through the general path."""
    params, pl, x = _cases.ms_case(pkg.synth, 3, seed=13, N=16000, wmin=0.004, wmax=0.02)
    Nmax = int(pl[0])
    params[0] = 1e-27
    params[1] = 3e24
    params[2] = 0.0          # a zero-height mode contributes exactly nothing
    rc, M = oracle.call_model(3, params, pl, x)
    assert rc == 0
    with _ctx(pkg, 3, params, pl, x, M) as ctx:
        Mg = ctx.model(params)
        assert np.all(np.isfinite(Mg))
        assert np.max(np.abs(Mg - M) / np.abs(M)) < RTOL
        L, st = ctx.eval(params[None, :])
        Lr = oracle.call_likelihood(M, M, 1.0, 1.0)
        assert abs(L[0, 0] - Lr) / abs(Lr) < RTOL


def test_full_size_properties(pkg):
    """BASELINE config C2 size (250k bins, 80 modes, 10 chains): size-independent properties only --
    tempering linearity, chain permutation equivariance, shard additivity, determinism."""
    import torch
    rng = np.random.default_rng(1)
    params, pl = pkg.synth.classic_params(rng)
    N = 250000
    x = pkg.synth.freq_axis(N, 500.0)
    T = pkg.synth.tcoefs(10, 1.7)
    with _ctx(pkg, 3, params, pl, x, np.ones(N), 1) as c0:
        M = c0.model(params)
    y = pkg.synth.chi2_2dof_spectrum(rng, M)
    P = pkg.synth.perturb_chains(rng, params, pl, 10)
    with _ctx(pkg, 3, params, pl, x, y, 10, T) as ctx:
        L, st = ctx.eval(P)
        assert (st == 0).all() and np.all(np.isfinite(L))
        L2, _ = ctx.eval(P)
        assert np.array_equal(L, L2)
        # permuting chains permutes S = -L*T
        perm = rng.permutation(10)
        Lp, _ = ctx.eval(P[perm])
        assert np.allclose(Lp[0] * T, (L[0] * T)[perm], rtol=1e-15, atol=0)
        # the truth beats a strongly perturbed model on average (sanity of the statistic)
        Pbad = P.copy(); Pbad[:, :20] *= 3.0
        Lb, _ = ctx.eval(Pbad)
        assert np.all(Lb[0] * T < L[0] * T)
    # shard additivity at full size
    halves = [(0, 124928), (124928, N)]
    S = np.zeros(10)
    for lo, hi in halves:
        with pkg.Context(pkg.Star.shard(3, pl, len(params), x, y, lo, hi), 10, T) as cs:
            dP = torch.tensor(cs.pack_params(P), device="cuda")
            dS = torch.zeros(10, dtype=torch.float64, device="cuda")
            cs.eval_device(dP.data_ptr(), dS.data_ptr(), raw_sum=True)
            cs.sync()
            S += dS.cpu().numpy()
    assert np.allclose(-S / T, L[0], rtol=1e-13, atol=0)


def test_chi_square_likelihood(pkg, oracle):
    """likelihood_chi_square (likelihoods.cpp:31-40, likelihoods_ctrl.list id 1): -sum((y-M)^2/sigma_y^2)/2, tempered like
    model_def.cpp:405; a missing sigma_y means ones (config.cpp:367-374)."""
    for model_id, seed in ((3, 21), (23, 22)):
        params, pl, x = _cases.ms_case(pkg.synth, model_id, seed=seed, N=25000, asym=0.0 if seed % 2 else 8.0)
        rc, M = oracle.call_model(model_id, params, pl, x)
        assert rc == 0
        rng = np.random.default_rng(seed)
        sigma = rng.uniform(0.3, 2.0, len(x)) * np.sqrt(M)
        y = M + sigma * rng.standard_normal(len(x))
        P = pkg.synth.perturb_chains(rng, params, pl, 3)
        T = pkg.synth.tcoefs(3, 1.7)
        for sig in (sigma, None):
            rc, L_ref = oracle.eval_chains_chi_square(model_id, P, pl, x, y, sigma if sig is not None else np.ones(len(x)), T)
            assert rc == 0
            star = pkg.Star(model_id, pl, len(params), x, y, sigma_y=sig)
            with pkg.Context(star, 3, T, likelihood_id=1) as ctx:
                L, st = ctx.eval(P)
                assert (st == 0).all()
                assert np.max(np.abs(L[0] - L_ref) / np.abs(L_ref)) < RTOL
                L2, _ = ctx.eval(P)
                assert np.array_equal(L, L2)


def test_tiny_and_large_spectra(pkg, oracle):
    """Size extremes: the smallest spectra the ABI accepts (2 and 3 bins) and a 3-million-bin spectrum (1954 tiles per chain)."""
    rng = np.random.default_rng(31)
    params, pl = pkg.synth.classic_params(rng, Nmax=3, lmax=2, f0=1000.0, dnu=60.0, trunc_c=10.0)
    for N in (2, 3, 5):
        x = pkg.synth.freq_axis(N, 1059.0, 0.25)
        rc, M = oracle.call_model(3, params, pl, x)
        assert rc == 0
        y = M * np.linspace(0.5, 1.5, N)
        rc, L_ref = oracle.eval_chains(3, params[None, :], pl, x, y, [1.0])
        with _ctx(pkg, 3, params, pl, x, y) as ctx:
            Mg = ctx.model(params)
            assert np.max(np.abs(Mg - M) / np.abs(M)) < RTOL
            L, st = ctx.eval(params[None, :])
            assert st[0, 0] == 0 and abs(L[0, 0] - L_ref[0]) / abs(L_ref[0]) < RTOL
    N = 3000000
    x = pkg.synth.freq_axis(N, 5.0, 0.003)
    params, pl = pkg.synth.classic_params(rng, Nmax=4, lmax=2, f0=3000.0, dnu=110.0, trunc_c=30.0)
    rc, M, tr = oracle.call_model(3, params, pl, x, trace=True)
    assert rc == 0
    y = pkg.synth.chi2_2dof_spectrum(rng, M)
    P = pkg.synth.perturb_chains(rng, params, pl, 2)
    T = [1.0, 1.7]
    rc, L_ref = oracle.eval_chains(3, P, pl, x, y, T)
    with _ctx(pkg, 3, params, pl, x, y, 2, T) as ctx:
        rcw, wl, w0, w1 = ctx.windows(params)
        assert np.array_equal(w0, tr[1]) and np.array_equal(w1, tr[2])
        L, st = ctx.eval(P)
        assert (st == 0).all()
        assert np.max(np.abs(L[0] - L_ref) / np.abs(L_ref)) < RTOL


@pytest.mark.parametrize("asym", [0.0, 10.0])
def test_c2_fullsize_against_reference_golden(pkg, oracle, asym):
    """BASELINE config C2 at FULL size (250k bins, 80 modes, 10 chains) against the log-likelihoods of the reference's own
    model + likelihood functions for the same seeded inputs (tests/golden/reference_c2_fullsize.json)."""
    import importlib.util
    import json
    import os
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_golden_c2_fullsize", os.path.join(gdir, "make_golden_c2_fullsize.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    gold = json.load(open(os.path.join(gdir, "reference_c2_fullsize.json")))["asym_%g" % asym]
    params, pl, x, y, P, T = mod.c2_inputs(pkg.synth, oracle, asym)
    assert abs(y.sum() - gold["y_sum"]) <= 1e-13 * abs(gold["y_sum"])
    with _ctx(pkg, 3, params, pl, x, y, 10, T) as ctx:
        L, st = ctx.eval(P)
        assert (st == 0).all()
        Lr = np.array(gold["logL_reference"])
        assert np.max(np.abs(L[0] - Lr) / np.abs(Lr)) < RTOL
        M0 = ctx.model(P[0])
        assert abs(M0.sum() - gold["model0_sum"]) <= 1e-11 * abs(gold["model0_sum"])
        for i, v in gold["model0_at"].items():
            assert abs(M0[int(i)] - v) <= RTOL * abs(v)


def test_device_entry_graph_cache_rotation(pkg, oracle):
    """tamcmc_gpu_eval_device keeps one CUDA graph per (params, active, out) pointer set (4 entries): rotating through more
    buffer sets than that must evict and re-capture without changing results; a caller stream is honoured."""
    import torch
    params, pl, x = _cases.ms_case(pkg.synth, 3, seed=17, N=9000)
    rc, M = oracle.call_model(3, params, pl, x)
    rng = np.random.default_rng(2)
    y = pkg.synth.chi2_2dof_spectrum(rng, M)
    T = pkg.synth.tcoefs(3, 1.7)
    sets = [pkg.synth.perturb_chains(rng, params, pl, 3) for _ in range(6)]
    refs = [oracle.eval_chains(3, P, pl, x, y, T)[1] for P in sets]
    stream = torch.cuda.Stream()
    with _ctx(pkg, 3, params, pl, x, y, 3, T) as ctx:
        dP = [torch.tensor(ctx.pack_params(P), device="cuda") for P in sets]
        dL = [torch.zeros(3, dtype=torch.float64, device="cuda") for _ in sets]
        for rnd in range(3):
            for k in range(6):
                with torch.cuda.stream(stream):
                    ctx.eval_device(dP[k].data_ptr(), dL[k].data_ptr(), stream=stream.cuda_stream)
            stream.synchronize()
            for k in range(6):
                got = dL[k].cpu().numpy()
                assert np.max(np.abs(got - refs[k]) / np.abs(refs[k])) < RTOL
                dL[k].zero_()
        # host entry interleaved with the device entry on the same context
        L, st = ctx.eval(sets[2])
        assert np.max(np.abs(L[0] - refs[2]) / np.abs(refs[2])) < RTOL


def test_device_parallel_tempering_swap(pkg):
    """tamcmc_gpu_pt_swap_device against the host rule of MALA::parallel_tempering (MALA.cpp:397-461) as restated in
    host/mcmc_driver.hpp: tempered log-likelihoods, adjacent chains, rows and priors exchanged only when accepted."""
    import torch
    params, pl, x = _cases.ms_case(pkg.synth, 3, seed=5, N=6000)
    rng = np.random.default_rng(8)
    Nch = 5
    T = pkg.synth.tcoefs(Nch, 1.7)
    stars = [pkg.Star(3, pl, len(params), x, np.ones_like(x)) for _ in range(2)]
    with pkg.Context(stars, Nch, T) as ctx:
        stride = ctx.params_stride
        P = rng.standard_normal((2, Nch, stride))
        for trial in range(40):
            L = -1e4 * (1 + 0.001 * rng.standard_normal((2, Nch))) / T          # tempered values close enough for both outcomes
            if trial == 7:
                L[1, 2] = np.nan
            pr = rng.standard_normal((2, Nch))
            star, A, u = int(rng.integers(0, 2)), int(rng.integers(0, Nch - 1)), float(rng.random())
            dP, dL, dpr = torch.tensor(P, device="cuda"), torch.tensor(L, device="cuda"), torch.tensor(pr, device="cuda")
            flag = torch.full((1,), -1, dtype=torch.int32, device="cuda")
            ctx.pt_swap_device(A, u, dP.data_ptr(), dL.data_ptr(), dpr.data_ptr(), flag.data_ptr(), star=star)
            ctx.sync()
            B = A + 1
            LA, LB = L[star, A], L[star, B]
            LA_TB, LB_TA = LA * T[A] / T[B], LB * T[B] / T[A]
            with np.errstate(over="ignore", invalid="ignore"):
                r = min(1.0, np.exp(LA_TB + LB_TA - LA - LB)) if np.isfinite(LA + LB) else np.nan
            want = bool(u <= r)
            assert int(flag[0]) == int(want)
            P2, L2, pr2 = P.copy(), L.copy(), pr.copy()
            if want:
                P2[star, [A, B]] = P[star, [B, A]]
                pr2[star, [A, B]] = pr[star, [B, A]]
                L2[star, A], L2[star, B] = LB_TA, LA_TB
            assert np.array_equal(dP.cpu().numpy(), P2)
            assert np.array_equal(dpr.cpu().numpy(), pr2)
            assert np.allclose(dL.cpu().numpy(), L2, rtol=1e-15, atol=0, equal_nan=True)
        with pytest.raises(pkg.TamcmcError):
            ctx.pt_swap_device(Nch - 1, 0.5, dP.data_ptr(), dL.data_ptr())      # no chain A+1


@pytest.mark.parametrize("asym", [0.0, 25.0])
def test_far_field_folding_against_per_bin_merge(pkg, oracle, monkeypatch, asym):
    """Modes far from a tile are folded into the tile's polynomial (whittle.cu, producer_loop).  The same library with
    TAMCMC_GPU_FAR_RATIO=0 merges every component per bin like the reference sums them (build_lorentzian.cpp:131-161):
    both must agree with the oracle to 1e-10 and with each other to 1e-12, at a ratio below the default as well."""
    params, pl, x = _cases.ms_case(pkg.synth, 3, seed=11, N=61000, asym=asym)
    rc, M = oracle.call_model(3, params, pl, x)[:2]
    assert rc == 0
    rng = np.random.default_rng(5)
    y = pkg.synth.chi2_2dof_spectrum(rng, M)
    P = pkg.synth.perturb_chains(rng, params, pl, 3)
    T = pkg.synth.tcoefs(3, 1.7)
    rc, L_ref = oracle.eval_chains(3, P, pl, x, y, T)
    assert rc == 0
    res = {}
    for ratio in ("0", None, "6"):
        if ratio is None:
            monkeypatch.delenv("TAMCMC_GPU_FAR_RATIO", raising=False)
        else:
            monkeypatch.setenv("TAMCMC_GPU_FAR_RATIO", ratio)
        with _ctx(pkg, 3, params, pl, x, y, 3, T) as ctx:
            Mg = ctx.model(params)
            L, st = ctx.eval(P)
            assert (st == 0).all()
        assert np.max(np.abs(Mg - M) / np.abs(M)) < RTOL
        assert np.max(np.abs(L[0] - L_ref) / np.abs(L_ref)) < RTOL
        res[ratio] = (Mg, L[0])
    for ratio in (None, "6"):
        assert np.max(np.abs(res[ratio][0] - res["0"][0]) / np.abs(res["0"][0])) < 1e-11
        assert np.max(np.abs(res[ratio][1] - res["0"][1]) / np.abs(res["0"][1])) < 1e-12
    # the folding is not a no-op on this case: the two paths round differently somewhere
    assert not np.array_equal(res[None][0], res["0"][0])


@pytest.mark.parametrize("asym,model", [(0.0, 3), (25.0, 3), (0.0, 23)])
def test_tiles_schedule_matches_ring_and_oracle(pkg, oracle, monkeypatch, asym, model):
    """TAMCMC_GPU_KERNEL=tiles selects the all-warps-on-one-tile schedule of the fused kernel (whittle_tiles.cu: persistent
    CTAs, mode tables prefetched by TMA one item ahead) instead of the producer / consumer ring (whittle.cu).  Same expander
    output, same per-tile partial sums: model spectrum and logL agree with the oracle to 1e-10 and with the ring to 1e-12,
    with and without the far-field folding; 61 000 bins = 40 tiles, so a CTA walks several items."""
    params, pl, x = _cases.ms_case(pkg.synth, model, seed=21, N=61000, asym=asym)
    rc, M = oracle.call_model(model, params, pl, x)[:2]
    assert rc == 0
    rng = np.random.default_rng(6)
    y = pkg.synth.chi2_2dof_spectrum(rng, M)
    P = pkg.synth.perturb_chains(rng, params, pl, 4)
    T = pkg.synth.tcoefs(4, 1.7)
    rc, L_ref = oracle.eval_chains(model, P, pl, x, y, T)
    assert rc == 0
    res = {}
    for kernel in ("ring", "tiles"):
        for ratio in (None, "0"):
            monkeypatch.setenv("TAMCMC_GPU_KERNEL", kernel)
            if ratio is None:
                monkeypatch.delenv("TAMCMC_GPU_FAR_RATIO", raising=False)
            else:
                monkeypatch.setenv("TAMCMC_GPU_FAR_RATIO", ratio)
            with _ctx(pkg, model, params, pl, x, y, 4, T) as ctx:
                Mg = ctx.model(params)
                L, st = ctx.eval(P)
                L2, _ = ctx.eval(P)
                assert (st == 0).all()
                assert np.array_equal(L, L2)          # bitwise reproducible
            assert np.max(np.abs(Mg - M) / np.abs(M)) < RTOL
            assert np.max(np.abs(L[0] - L_ref) / np.abs(L_ref)) < RTOL
            res[(kernel, ratio)] = (Mg, L[0])
    for ratio in (None, "0"):
        a, b = res[("ring", ratio)], res[("tiles", ratio)]
        assert np.max(np.abs(a[0] - b[0]) / np.abs(a[0])) < 1e-12
        assert np.max(np.abs(a[1] - b[1]) / np.abs(a[1])) < 1e-12


def test_rows_built_in_the_staging_block(pkg, oracle):
    """tamcmc_gpu_params_staging: rows written into the context's own pinned block and passed back as `params` give the same
    results as rows passed from the caller's memory (the call only skips its host-side copy)."""
    params, pl, x = _cases.ms_case(pkg.synth, 3, seed=4, N=20000)
    rc, M = oracle.call_model(3, params, pl, x)[:2]
    rng = np.random.default_rng(2)
    y = pkg.synth.chi2_2dof_spectrum(rng, M)
    P = pkg.synth.perturb_chains(rng, params, pl, 3)
    T = pkg.synth.tcoefs(3, 1.5)
    with _ctx(pkg, 3, params, pl, x, y, 3, T) as ctx:
        L0, st0 = ctx.eval(P)
        S = ctx.params_staging()
        assert S.shape == (1, 3, ctx.params_stride)
        S[...] = 0.0
        S[0, :, :P.shape[1]] = P
        L1, st1 = ctx.eval(S)
        assert np.array_equal(L0, L1) and np.array_equal(st0, st1)
        S[0, 1, :P.shape[1]] = P[2]                     # rewritten in place: the next call sees it
        L2, _ = ctx.eval(S)
        assert L2[0, 1] * T[1] == pytest.approx(L0[0, 2] * T[2], rel=1e-12)
