"""Synthetic Kepler-like inputs for the hot path (fixed seeds; numpy only).

Parameter-vector layouts follow the reference's `.model` reader:
plength = [Nmax, lmax, Nfl0, Nfl1, Nfl2, Nfl3, Nsplit, Nwidth, Nnoise, Ninc, Ncfg]
(tamcmc/sources/io_ms_global.cpp:1315-1325) and
params  = [H | V_l | fl0 | fl1 | fl2 | fl3 | split | W | noise | inc | cfg]
(tamcmc/sources/io_ms_global.cpp:1329-1398).  The recipes are the ones SURVEY.md 8(d)
names; `make_params_aj_model` mirrors the input recipe of the reference's own unit test
(test/lorentzian_test/unit_tests/test_build_l_mode.cpp:769-874) with a fixed seed.
"""
import numpy as np

# Kepler long-cadence-free resolution used by the reference tests:
# test/lorentzian_test/unit_tests/test_build_l_mode.cpp:107
RESOL_4YR = 1e6 / (4.0 * 365.0 * 86400.0)

MODEL_CLASSIC = 3
MODEL_A1L_ETAA3 = 6
MODEL_LOCAL_BASIC = 11
MODEL_CLASSIC_V2 = 12
MODEL_CLASSIC_V3 = 13
MODEL_AJALM = 21
MODEL_AJ = 23
MODEL_MODE_TABLE = 1000      # include/tamcmc_gpu.h: TAMCMC_MODEL_MODE_TABLE
MT_HEADER, MT_STRIDE = 4, 20


def freq_axis(N, x0, step=RESOL_4YR):
    return x0 + step * np.arange(N, dtype=np.float64)


def tcoefs(Nchains, lam):
    """T_m = lambda^m (Config/default/config_default.cfg: Tcoef)."""
    return lam ** np.arange(Nchains, dtype=np.float64)


def ms_global_modes(rng, Nmax=20, lmax=3, f0=650.0, dnu=85.0, scatter=0.01):
    """Central frequencies fl[l][n] of a main-sequence comb."""
    n = np.arange(Nmax)
    off = {0: 0.0, 1: 0.5 * dnu - 2.0, 2: -6.0, 3: 0.5 * dnu - 14.0}
    fl = []
    for l in range(lmax + 1):
        fl.append(f0 + dnu * n + off[l] + rng.uniform(-scatter * dnu, scatter * dnu, Nmax))
    return fl


def classic_params(rng, Nmax=20, lmax=3, f0=650.0, dnu=85.0, asym=0.0, inc=45.0, a1=1.0, a3=0.01,
                   trunc_c=30.0, do_amp=0, noise=None, wmin=1.0, wmax=8.0):
    """model_MS_Global_a1etaa3_HarveyLike_Classic (models.cpp:1943): Nsplit=6
    [a1, eta, a3, magb, magalfa, asym], Ninc=1, Ncfg=2."""
    if noise is None:
        noise = [0.0, 0.0, 1.0, 1.0, 100.0, 2.0, 0.5, 10.0, 2.0, 0.1]
    fl = ms_global_modes(rng, Nmax, lmax, f0, dnu)
    n = np.arange(Nmax)
    H = rng.uniform(10.0, 20.0, Nmax)
    W = wmin + (wmax - wmin) * (0.5 - 0.5 * np.cos(np.pi * n / max(Nmax - 1, 1))) + rng.uniform(0, 0.05, Nmax)
    V = np.array([1.5, 0.53, 0.08])[:lmax]
    split = np.array([a1, 0.0, a3, 0.0, 0.0, asym])
    params = np.concatenate([H, V] + fl + [split, W, np.asarray(noise, float), [inc], [trunc_c, float(do_amp)]])
    plength = np.array([Nmax, lmax] + [Nmax if l <= lmax else 0 for l in range(4)] +
                       [len(split), Nmax, len(noise), 1, 2], dtype=np.int32)
    return params, plength


def aj_params(rng, Nmax=20, lmax=3, f0=650.0, dnu=85.0, asym=0.0, inc=45.0, a1=1.0, trunc_c=30.0,
              do_amp=0, noise=None, eta_switch=1.0, wmin=1.0, wmax=8.0):
    """model_MS_Global_aj_HarveyLike (models.cpp:1195): Nsplit=14 =
    [a1_0,a1_1, a2_0,a2_1, ..., a6_0,a6_1, eta_switch, asym]."""
    if noise is None:
        noise = [0.0, 0.0, 1.0, 1.0, 100.0, 2.0, 0.5, 10.0, 2.0, 0.1]
    fl = ms_global_modes(rng, Nmax, lmax, f0, dnu)
    n = np.arange(Nmax)
    H = rng.uniform(10.0, 20.0, Nmax)
    W = wmin + (wmax - wmin) * (0.5 - 0.5 * np.cos(np.pi * n / max(Nmax - 1, 1))) + rng.uniform(0, 0.05, Nmax)
    V = np.array([1.5, 0.53, 0.08])[:lmax]
    aj = np.zeros(12)
    aj[0] = a1
    aj[1] = 0.02
    aj[2] = rng.uniform(-0.1 * a1, 0.1 * a1)
    aj[4] = rng.uniform(-0.025 * a1, 0.025 * a1)
    aj[6] = rng.uniform(-0.025 * a1, 0.025 * a1)
    aj[8] = rng.uniform(-0.01 * a1, 0.01 * a1)
    aj[10] = rng.uniform(-0.005 * a1, 0.005 * a1)
    split = np.concatenate([aj, [eta_switch, asym]])
    params = np.concatenate([H, V] + fl + [split, W, np.asarray(noise, float), [inc], [trunc_c, float(do_amp)]])
    plength = np.array([Nmax, lmax] + [Nmax if l <= lmax else 0 for l in range(4)] +
                       [len(split), Nmax, len(noise), 1, 2], dtype=np.int32)
    return params, plength


def ajalm_params(rng, Nmax=11, lmax=2, f0=2100.0, dnu=103.0, asym=0.0, inc=60.0, a1=1.2, trunc_c=30.0, do_amp=0, noise=None,
                 eta_switch=1.0, epsilon=5e-3, theta0=50.0, delta=20.0, decompose_Alm=1, filter_code=0, wmin=1.0, wmax=5.0):
    """model_MS_Global_ajAlm_HarveyLike (models.cpp:1411): Nsplit=12 =
    [a1_0,a1_1, a3_0,a3_1, a5_0,a5_1, eps_0,eps_1, theta0(deg), delta(deg), eta_switch, asym], Ncfg=4 =
    [trunc_c, do_amp, decompose_Alm, filter_code]."""
    if noise is None:
        noise = [0.0, 0.0, 1.0, 1.0, 100.0, 2.0, 0.5, 10.0, 2.0, 0.1]
    fl = ms_global_modes(rng, Nmax, lmax, f0, dnu)
    n = np.arange(Nmax)
    H = rng.uniform(10.0, 20.0, Nmax)
    W = wmin + (wmax - wmin) * (0.5 - 0.5 * np.cos(np.pi * n / max(Nmax - 1, 1))) + rng.uniform(0, 0.05, Nmax)
    V = np.array([1.5, 0.53, 0.08])[:lmax]
    split = np.array([a1, 0.01, rng.uniform(-0.02, 0.02), 0.0, rng.uniform(-0.005, 0.005), 0.0, epsilon, 1e-4, theta0, delta, eta_switch, asym])
    cfg = [trunc_c, float(do_amp), float(decompose_Alm), float(filter_code)]
    params = np.concatenate([H, V] + fl + [split, W, np.asarray(noise, float), [inc], cfg])
    plength = np.array([Nmax, lmax] + [Nmax if l <= lmax else 0 for l in range(4)] + [len(split), Nmax, len(noise), 1, len(cfg)],
                       dtype=np.int32)
    return params, plength


def make_params_aj_model(rng, lmax, Nfreqs, Dnu, epsilon, d0l, asym_on=None):
    """The reference unit test's input recipe for model_MS_Global_aj_HarveyLike
    (test_build_l_mode.cpp:769-874), seeded.  Note the reference's `el/2` is an
    INTEGER division (el is int)."""
    trunc = 50.0
    fl = []
    for el in range(lmax + 1):
        for en in range(Nfreqs):
            sc = rng.uniform(-Dnu / 100, Dnu / 100)
            fl.append((en + epsilon + el // 2) * Dnu + d0l * el * (el + 1) + sc)
    fl = np.array(fl)
    aj = np.zeros(13)
    aj[0] = rng.uniform(0.1, 5)
    aj[2] = rng.uniform(-0.1 * aj[0], 0.1 * aj[0])
    aj[4] = rng.uniform(-0.025 * aj[0], 0.025 * aj[0])
    aj[6] = rng.uniform(-0.025 * aj[0], 0.025 * aj[0])
    aj[8] = rng.uniform(-0.01 * aj[0], 0.01 * aj[0])
    aj[10] = rng.uniform(-0.005 * aj[0], 0.005 * aj[0])
    aj[12] = 0.0
    if asym_on is None:
        asym_on = bool(rng.integers(0, 2))
    asym = rng.uniform(-100, 100.0) if asym_on else 0.0
    vis = np.array([1.5, 0.53, 0.07])[:lmax]
    H = rng.uniform(10, 20, Nfreqs)
    W = rng.uniform(0.5, 2, Nfreqs)
    noise = np.array([0, 1, 1, 0, 1, 1, 0.1], float)
    inc = rng.uniform(0.0, 90.0)
    params = np.concatenate([H, vis, fl, aj, [asym], W, noise, [inc], [trunc, 0.0], [0.0]])
    Nfl = [Nfreqs if l <= lmax else 0 for l in range(4)]
    plength = np.array([Nfreqs, lmax] + Nfl + [len(aj) + 1, Nfreqs, len(noise), 1, 2], dtype=np.int32)
    return params, plength


def perturb_chains(rng, params, plength, Nchains, rel=0.01):
    """Chain parameter vectors = truth + small perturbations of the fitted quantities
    (heights, frequencies, widths, noise); configuration slots are left untouched."""
    Nmax, lmax = int(plength[0]), int(plength[1])
    Nf = int(plength[2] + plength[3] + plength[4] + plength[5])
    Nsplit, Nwidth, Nnoise, Ninc = (int(plength[i]) for i in (6, 7, 8, 9))
    P = np.tile(params, (Nchains, 1))
    o = 0
    P[:, o:o + Nmax] *= 1 + rel * rng.standard_normal((Nchains, Nmax)); o += Nmax
    P[:, o:o + lmax] *= 1 + rel * rng.standard_normal((Nchains, lmax)); o += lmax
    P[:, o:o + Nf] += 0.1 * rel * 85.0 * rng.standard_normal((Nchains, Nf)); o += Nf
    o += Nsplit
    P[:, o:o + Nwidth] *= 1 + rel * rng.standard_normal((Nchains, Nwidth)); o += Nwidth
    nz = params[o:o + Nnoise] != 0
    P[:, o:o + Nnoise][:, nz] *= 1 + rel * rng.standard_normal((Nchains, int(nz.sum()))); o += Nnoise
    P[:, o:o + Ninc] += rel * 10 * rng.standard_normal((Nchains, Ninc))
    P[0] = params
    return np.ascontiguousarray(P)


def chi2_2dof_spectrum(rng, model):
    """y_i = M_i * E_i, E ~ Exp(1): a chi^2 with 2 d.o.f. power spectrum around the model."""
    return model * rng.exponential(1.0, size=model.shape)


def mode_table_plength(capacity, Nnoise, step_mode=0):
    """plength of the generic mode table (include/tamcmc_gpu.h)."""
    pl = np.zeros(11, dtype=np.int32)
    pl[0], pl[1], pl[8] = capacity, step_mode, Nnoise
    return pl


def mode_table_nparams(capacity, Nnoise):
    return MT_HEADER + Nnoise + MT_STRIDE * capacity


def mode_table_row(capacity, inclination, trunc_c, asym, noise, modes, extra=None):
    """Pack one chain's resolved modes into a mode-table parameter row.
    modes: array [nmodes, >=11] with columns l, fc, H, W, a1..a6, eta0 (the first 11 columns of the reference's own
    `mode_params` table, models.cpp:4941, with eta0 unscaled); extra: optional [nmodes, 7] per-m frequency shifts."""
    modes = np.asarray(modes, dtype=np.float64)
    noise = np.asarray(noise, dtype=np.float64)
    n = len(modes)
    if n > capacity:
        raise ValueError("more modes than the table capacity")
    row = np.zeros(mode_table_nparams(capacity, len(noise)))
    row[0:4] = [n, inclination, trunc_c, asym]
    row[4:4 + len(noise)] = noise
    rec = np.zeros((capacity, MT_STRIDE))
    rec[:n, :11] = modes[:, :11]
    if extra is not None:
        rec[:n, 11:18] = np.asarray(extra, dtype=np.float64)
    row[4 + len(noise):] = rec.ravel()
    return row


# ---- Gaussian-envelope models (ids 0 and 1; models.cpp:5728-5797, 5674-5725): fixed parameter positions, no plength ----
ENVELOPE_PLENGTH = [0] * 11


def kallinger_gaussian_params(rng=None, numax=100.0, jitter=0.0):
    """[k_a, s_a, k_b0, s_b0, c0, a1, a2, k1, s1, c1, k2, s2, c2, N0, Amax, numax, sigma, mu_numax] with the scaling
    relations of Kallinger+2014 Table 2 (amplitudes in ppm, frequencies in microHz)."""
    p = np.array([3335.0, -0.564, 0.317, 0.970, 4.0, 3382.0 * numax ** -0.609, 3382.0 * numax ** -0.609,
                  0.317, 0.970, 4.0, 0.948, 0.992, 4.0, 5.0, 800.0, numax, 0.12 * numax, 0.5])
    if rng is not None and jitter > 0:
        p = p * (1.0 + jitter * rng.standard_normal(p.size))
    return p


def harvey_gaussian_params(rng=None, numax=120.0, jitter=0.0):
    """[H1, tc1, p1, H2, tc2, p2, B0, Hgauss, nu_gauss, sigma]"""
    p = np.array([50.0, 30.0, 2.0, 10.0, 5.0, 3.0, 2.0, 30.0, numax, 0.12 * numax])
    if rng is not None and jitter > 0:
        p = p * (1.0 + jitter * rng.standard_normal(p.size))
    return p
