"""Device red-giant expander (tamcmc_gpu_rgb_expand: csrc/rgb_device.cu, rgb_solver.cuh, dd_math.cuh) -- the pair loop of the ARMM
mixed-mode solver (external/ARMM/solver_mm.cpp:326-449, 558-573) and the zeta normalisation (external/ARMM/bump_DP.cpp:126-163) on the
GPU, all chains of a step in one call.

CPU (no GPU): the double-double tan / atan the device uses are correctly rounded (against mpmath) and equal glibc's wherever glibc's are;
the segment decomposition run on the host reproduces tamcmc_host_expand_rgb_v4 -- which is pinned bit for bit on the reference's own
functions (tests/test_rgb_expander.py) -- bit for bit with glibc's tan / atan, and with the device's ones up to the stated 1 ulp.
GPU: rows of tamcmc_gpu_rgb_expand against the host expander's (frequencies identical or 1 ulp, everything else 1e-12), model spectrum and
logL from reference parameter vectors against the reference's own output at 1e-10."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden", "reference_rgb_vectors.npz")


def _variants(G, n_per_case, seed=7):
    """the fixture's reference parameter vectors + perturbed copies (delta0l, DPl, alpha_g, q, l=0 frequencies)"""
    rng = np.random.default_rng(seed)
    for case in range(int(G["ncases"])):
        for rep in range(n_per_case):
            params, pl = G["params%d" % case].copy(), G["plength%d" % case]
            if rep:
                Nmax, lmax, Nfl0 = int(pl[0]), int(pl[1]), int(pl[2])
                o = Nmax + lmax + Nfl0
                params[o] += rng.normal() * 0.02
                params[o + 1] *= 1 + rng.normal() * 0.01
                params[o + 2] += rng.normal() * 0.05
                params[o + 3] *= 1 + rng.normal() * 0.1
                params[Nmax + lmax:Nmax + lmax + Nfl0] += rng.normal(size=Nfl0) * 0.03
            yield case, rep, params, pl


def test_dd_tan_atan_are_correctly_rounded(tmp_path):
    """tan_cr / atan_cr (csrc/dd_math.cuh, compiled for the host) against mpmath at 200 bits, and against glibc: they may only differ
    where glibc is not correctly rounded."""
    mp = pytest.importorskip("mpmath")
    src = tmp_path / "t.cpp"
    src.write_text(r'''
#include "dd_math.cuh"
#include <cstdio>
#include <cstdlib>
#include <random>
int main(int argc, char** argv) {
    std::mt19937_64 rng(2024);
    const long N = atol(argv[1]);
    std::uniform_real_distribution<double> U(-400.0, 400.0), L(-17.0, 17.0);
    long bt = 0, ba = 0;
    for (long i = 0; i < N; i++) {
        const double x = U(rng), a = std::tan(x), b = tamcmc_dd::tan_cr(x);
        double t = std::pow(10.0, L(rng)); if (i & 1) t = -t;
        const double c = std::atan(t), d = tamcmc_dd::atan_cr(t);
        if (a != b) bt++;
        if (c != d) ba++;
        if (i < 3000 || a != b) printf("tan %a %a\n", x, b);
        if (i < 3000 || c != d) printf("atan %a %a\n", t, d);
    }
    printf("count %ld %ld\n", bt, ba);
}''')
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "tamcmc-c_b200", "csrc"), str(src), "-o", str(exe)])
    N = 400000
    out = subprocess.check_output([str(exe), str(N)], text=True).split("\n")
    mp.mp.prec = 200
    checked = 0
    for line in out:
        f = line.split()
        if len(f) == 3 and f[0] in ("tan", "atan"):
            x, v = float.fromhex(f[1]), float.fromhex(f[2])
            exact = mp.tan(mp.mpf(x)) if f[0] == "tan" else mp.atan(mp.mpf(x))
            err = abs(mp.mpf(v) - exact)
            assert err <= abs(mp.mpf(float(np.nextafter(v, np.inf))) - exact) and err <= abs(mp.mpf(float(np.nextafter(v, -np.inf))) - exact), (f[0], x)
            checked += 1
        elif len(f) == 3 and f[0] == "count":
            # glibc 2.39: tan is not correctly rounded for ~0.25 % of the arguments, atan for ~0.02 %
            assert int(f[1]) < 0.01 * N and int(f[2]) < 0.002 * N
    assert checked >= 6000


def test_dd_tan_atan_edge_cases(tmp_path):
    """tan_cr next to 570 zeros and poles of the tangent (distances 1e-1 ... 1e-11), tiny and large arguments; atan_cr from 1e-300 to 1e280:
    correctly rounded everywhere (mpmath, 400 bits)."""
    mp = pytest.importorskip("mpmath")
    exe = tmp_path / "edge"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "tamcmc-c_b200", "csrc"),
                           os.path.join(HERE, "cpp", "dd_edge_cases.cpp"), "-o", str(exe)])
    mp.mp.prec = 400
    n = 0
    for line in subprocess.check_output([str(exe)], text=True).splitlines():
        k, x, v = line.split()
        x, v = float.fromhex(x), float.fromhex(v)
        exact = mp.tan(mp.mpf(x)) if k == "tan" else mp.atan(mp.mpf(x))
        n += 1
        if v == 0 and exact == 0:
            continue
        err = abs(mp.mpf(v) - exact)
        assert err <= abs(mp.mpf(float(np.nextafter(v, np.inf))) - exact) and err <= abs(mp.mpf(float(np.nextafter(v, -np.inf))) - exact), (k, x)
    assert n > 5000


def test_long_double_local_grid_is_reproduced_exactly(tmp_path):
    """local_grid_ext (csrc/rgb_solver.cuh): the reference's x87 extended-precision local grid (solver_mm.cpp:402-410) from error-free
    transformations in double, against the real long double arithmetic on 3 million random (nu, resol, factor): range_min, range_max and
    the truncated point count (which flips between 799 and 800 with the last bits of the quotient) must all be identical."""
    exe = tmp_path / "lg"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "tamcmc-c_b200", "csrc"),
                           os.path.join(HERE, "cpp", "local_grid_ext_check.cpp"), "-o", str(exe)])
    out = subprocess.check_output([str(exe), "3000000"], text=True)
    f = out.split()
    assert "bad 0 notok 0" in out, out
    assert int(f[7].rstrip(",")) > 1000 and int(f[9].rstrip(")")) > 1000, out          # both truncations occur


def test_borderline_ratio_test_in_double_double(tmp_path):
    """ratio_dd (csrc/rgb_solver.cuh): g / p of a proposed solution in double-double against the reference's long double formula
    (solver_mm.cpp:172-180, 421-427) on 10^6 random inputs -- within the extended-precision rounding errors (1e-14 + 4e-14 / |p|), a quarter
    of the band inside which the device leaves the decision to the host."""
    exe = tmp_path / "ratio"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "tamcmc-c_b200", "csrc"),
                           os.path.join(HERE, "cpp", "ratio_dd_check.cpp"), "-o", str(exe)])
    out = subprocess.check_output([str(exe), "1000000"], text=True)
    assert " bad 0 " in out, out


@pytest.mark.parametrize("model_id", [25, 27])
def test_segment_decomposition_reproduces_the_host_solver(pkg, model_id):
    G = np.load(GOLD)
    step = G["x"][2] - G["x"][1]
    n = 0
    for case, rep, params, pl in _variants(G, 3):
        try:
            row, nm = pkg.expand_rgb_v4(model_id, params, pl, step, 120)
        except pkg.TamcmcError:
            continue
        r0, nm0, fl0 = pkg.expand_rgb_v4_emulated(model_id, params, pl, step, 120, exact_trig=0)
        assert fl0 == 0 and nm0 == nm and np.array_equal(r0, row), (case, rep)             # same operations, same library: same bits
        r1, nm1, fl1 = pkg.expand_rgb_v4_emulated(model_id, params, pl, step, 120, exact_trig=1)
        assert fl1 == 0 and nm1 == nm
        nn = int(pl[8])
        a, b = row[4 + nn:4 + nn + 20 * nm].reshape(nm, 20), r1[4 + nn:4 + nn + 20 * nm].reshape(nm, 20)
        assert np.all(np.abs(a[:, 1] - b[:, 1]) <= np.spacing(a[:, 1])), (case, rep)          # frequencies: identical or 1 ulp
        assert np.mean(a[:, 1] == b[:, 1]) > 0.95
        np.testing.assert_allclose(b, a, rtol=1e-12, atol=0)
        n += 1
    assert n >= 10


def _synthetic_giants(G, ntrials, seed=11):
    """other stars than the fixture: large separations of 3-18 microHz, period spacings of 60-95 s, couplings of 0.05-0.5, coarser and finer
    spectra, both solver entry points and the three bias settings -- built on the fixture's parameter layout"""
    rng = np.random.default_rng(seed)
    pl, base = G["plength0"], G["params0"]
    step = G["x"][2] - G["x"][1]
    Nmax, lmax, Nfl0, Nfl1, Nfl2, Nfl3 = [int(v) for v in pl[:6]]
    o0 = Nmax + lmax
    o1 = o0 + Nfl0
    o2 = o1 + Nfl1
    o3 = o2 + Nfl2
    for trial in range(ntrials):
        p = base.copy()
        Dnu = rng.uniform(3.0, 18.0)
        numax = (Dnu / 0.267) ** (1 / 0.764)
        n0 = int(numax / Dnu) - Nmax // 2
        fl0 = (n0 + np.arange(Nmax) + rng.uniform(0.8, 1.3)) * Dnu + rng.normal(size=Nmax) * 0.01 * Dnu
        p[o0:o0 + Nfl0] = fl0
        p[o1] = rng.normal() * 0.02 * Dnu / 9
        p[o1 + 1], p[o1 + 2], p[o1 + 3] = rng.uniform(60, 95), rng.uniform(0, 1), rng.uniform(0.05, 0.5)
        Nferr = int(p[-1])
        p[o1 + 8:o1 + 8 + Nferr] = np.linspace(fl0.min() - Dnu, fl0.max() + Dnu, Nferr)
        p[o1 + 8 + Nferr:o1 + 8 + 2 * Nferr] = rng.normal(size=Nferr) * 0.02
        if Nfl2 <= Nmax:
            p[o2:o2 + Nfl2] = fl0[:Nfl2] - 0.12 * Dnu
        if Nfl3 <= Nmax:
            p[o3:o3 + Nfl3] = fl0[:Nfl3] + 0.2 * Dnu
        p[-3], p[-2] = trial % 2, trial % 3
        yield trial, p, pl, step * rng.choice([1.0, 1.0, 0.5, 2.0])


def test_segment_decomposition_on_other_stars(pkg):
    G = np.load(GOLD)
    n = 0
    for trial, p, pl, stp in _synthetic_giants(G, 12):
        for model_id in (25, 27):
            try:
                row, nm = pkg.expand_rgb_v4(model_id, p, pl, stp, 400)
            except pkg.TamcmcError:
                continue
            for exact in (0, 1):
                try:
                    r, nm2, fl = pkg.expand_rgb_v4_emulated(model_id, p, pl, stp, 400, exact_trig=exact)
                except pkg.TamcmcError:
                    continue                          # flagged part-way (e.g. poles a few grid steps apart): the device path hands the chain to the host
                if fl:
                    continue
                assert nm2 == nm
                nn = int(pl[8])
                a, b = row[4 + nn:4 + nn + 20 * nm].reshape(nm, 20), r[4 + nn:4 + nn + 20 * nm].reshape(nm, 20)
                if exact == 0:
                    assert np.array_equal(r, row), trial
                else:
                    assert np.all(np.abs(a[:, 1] - b[:, 1]) <= np.spacing(a[:, 1])) and np.mean(a[:, 1] == b[:, 1]) > 0.95
                    np.testing.assert_allclose(b, a, rtol=1e-12, atol=0)
                n += 1
    assert n >= 30


def test_gpu_rgb_symbols_fail_loudly_without_a_device(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    G = np.load(GOLD)
    with pytest.raises(pkg.TamcmcError) as e:
        pkg.RgbExpander(25, G["plength0"], 0.01, 100, 4)
    assert e.value.status == pkg.ERR_CUDA


@pytest.mark.gpu
@pytest.mark.parametrize("model_id", [25, 27])
def test_gpu_rgb_expand_matches_host_expander(pkg, model_id):
    G = np.load(GOLD)
    step = G["x"][2] - G["x"][1]
    cases = [(c, r, p, pl) for c, r, p, pl in _variants(G, 5)]          # 20 chains: two groups in flight (>= 16)
    pl = cases[0][3]
    assert all(np.array_equal(pl, c[3]) for c in cases)
    P = np.stack([c[2] for c in cases])
    with pkg.RgbExpander(model_id, pl, step, 120, len(cases)) as rx:
        rows, nm, st, path = rx.expand(P)
        rows2, nm2, st2, path2 = rx.expand(P)
        assert np.array_equal(rows, rows2) and np.array_equal(nm, nm2)                     # deterministic (the candidate order is not, the set is)
    nn = int(pl[8])
    ndev = 0
    for i, (case, rep, params, _) in enumerate(cases):
        try:
            row, n = pkg.expand_rgb_v4(model_id, params, pl, step, 120)
        except pkg.TamcmcError as e:
            assert st[i] == e.status
            continue
        assert st[i] == 0 and nm[i] == n, (case, rep, st[i], nm[i], n)
        ndev += int(path[i] == 0)
        a, b = row[4 + nn:4 + nn + 20 * n].reshape(n, 20), rows[i, 4 + nn:4 + nn + 20 * n].reshape(n, 20)
        assert np.array_equal(a[:, 0], b[:, 0])
        assert np.all(np.abs(a[:, 1] - b[:, 1]) <= np.spacing(a[:, 1])), (case, rep)
        assert np.mean(a[:, 1] == b[:, 1]) > 0.95
        np.testing.assert_allclose(b, a, rtol=1e-12, atol=0)
        assert np.array_equal(rows[i, :4 + nn], row[:4 + nn])
    assert ndev >= len(cases) - 1                                                          # the device path is the one that ran


@pytest.mark.gpu
def test_gpu_rgb_expand_on_other_stars(pkg):
    """the synthetic giants of the CPU test through the kernels, chains with different numbers of modes in one call"""
    G = np.load(GOLD)
    cases = list(_synthetic_giants(G, 12))
    pl = cases[0][2]
    by_step = {}
    for trial, p, _, stp in cases:
        by_step.setdefault(float(stp), []).append(p)
    nn = int(pl[8])
    ndev = ntot = 0
    for stp, plist in by_step.items():
        P = np.stack(plist)
        with pkg.RgbExpander(25, pl, stp, 400, len(plist)) as rx:
            rows, nm, st, path = rx.expand(P)
        for i, p in enumerate(plist):
            try:
                row, n = pkg.expand_rgb_v4(25, p, pl, stp, 400)
            except pkg.TamcmcError as e:
                assert st[i] == e.status
                continue
            assert st[i] == 0 and nm[i] == n
            a, b = row[4 + nn:4 + nn + 20 * n].reshape(n, 20), rows[i, 4 + nn:4 + nn + 20 * n].reshape(n, 20)
            assert np.all(np.abs(a[:, 1] - b[:, 1]) <= np.spacing(a[:, 1]))
            np.testing.assert_allclose(b, a, rtol=1e-12, atol=0)
            ndev += int(path[i] == 0)
            ntot += 1
    assert ntot >= 10 and ndev >= ntot - 3


@pytest.mark.gpu
def test_gpu_rgb_expand_over_a_cloud_of_proposals(pkg):
    """What an MCMC run does: 120 parameter vectors scattered around the fixture's (every hyper-parameter of the mixed modes, the l=0
    frequencies, heights, widths), six calls of 20 chains.  Every frequency within 1 ulp of the host solver's, at least 99 % identical;
    the count of chains handed to the host solver is reported by the handle and stays small."""
    G = np.load(GOLD)
    step = G["x"][2] - G["x"][1]
    pl = G["plength0"]
    rng = np.random.default_rng(2024)
    Nmax, lmax, Nfl0 = int(pl[0]), int(pl[1]), int(pl[2])
    o = Nmax + lmax + Nfl0
    nn = int(pl[8])
    tot = same = 0
    with pkg.RgbExpander(25, pl, step, 140, 20) as rx:
        for call in range(6):
            base = G["params%d" % (call % 4)]
            P = np.tile(base, (20, 1))
            P[:, o] += rng.normal(size=20) * 0.03
            P[:, o + 1] *= 1 + rng.normal(size=20) * 0.01
            P[:, o + 2] = np.abs(P[:, o + 2] + rng.normal(size=20) * 0.05)
            P[:, o + 3] *= np.exp(rng.normal(size=20) * 0.1)
            P[:, Nmax + lmax:o] += rng.normal(size=(20, Nfl0)) * 0.02
            P[:, :Nmax] *= np.exp(rng.normal(size=(20, Nmax)) * 0.05)
            rows, nm, st, path = rx.expand(P)
            for i in range(20):
                try:
                    row, n = pkg.expand_rgb_v4(25, P[i], pl, step, 140)
                except pkg.TamcmcError as e:
                    assert st[i] == e.status
                    continue
                assert st[i] == 0 and nm[i] == n
                a, b = row[4 + nn:4 + nn + 20 * n].reshape(n, 20), rows[i, 4 + nn:4 + nn + 20 * n].reshape(n, 20)
                l1 = a[:, 0] == 1
                assert np.all(np.abs(a[:, 1] - b[:, 1]) <= np.spacing(a[:, 1]))
                np.testing.assert_allclose(b, a, rtol=1e-11, atol=0)
                tot += int(l1.sum())
                same += int((a[l1, 1] == b[l1, 1]).sum())
        n_setups, n_host = rx.counts()
    print("mixed-mode frequencies identical to the host solver: %d of %d; chains handed to the host solver: %d of %d" % (same, tot, n_host, n_setups))
    assert tot > 3000 and same >= 0.99 * tot
    assert n_setups == 120 and n_host <= 6


@pytest.mark.gpu
def test_gpu_rgb_params_to_logl_against_reference_model(pkg, oracle):
    """reference parameter vectors -> tamcmc_gpu_rgb_expand (rows in the context's staging block) -> tamcmc_gpu_eval, against the
    spectrum the reference's own model function returned for the same vectors (tests/golden/reference_rgb_vectors.npz)."""
    G = np.load(GOLD)
    x, y = G["x"], G["y"]
    step = x[2] - x[1]
    synth = pkg.synth
    n = int(G["ncases"])
    pl = G["plength0"]
    if not all(np.array_equal(pl, G["plength%d" % i]) for i in range(n)):
        pytest.skip("fixture cases differ in layout")
    P = np.stack([G["params%d" % i] for i in range(n)])
    cap, nn = 110, int(pl[8])
    T = np.ones(n)
    star = pkg.Star(synth.MODEL_MODE_TABLE, synth.mode_table_plength(cap, nn, 1), synth.mode_table_nparams(cap, nn), x, y)
    with pkg.Context(star, n, T) as ctx, pkg.RgbExpander(25, pl, step, cap, n) as rx:
        stage = ctx.params_staging()[0]
        rows, nm, st, path = rx.expand(P, rows_out=stage)
        assert (st == 0).all() and (path == 0).all()
        L, cs = ctx.eval(stage)
        assert (cs == 0).all()
        rows = np.array(stage[:n])                  # (tamcmc_gpu_model stages its own row in the same block)
        for i in range(n):
            M_ref = G["model%d" % i]
            M = ctx.model(rows[i])
            assert np.max(np.abs(M - M_ref) / np.abs(M_ref)) < 1e-10, i
            L_ref = oracle.call_likelihood(y, M_ref, 1.0, T[i])
            assert abs(L[0, i] - L_ref) <= 1e-10 * abs(L_ref)
