"""The Alm activity term from precomputed grids (what model_MS_Global_ajAlm_HarveyLike really uses, config C3):
grid files, GridMaker, and the bicubic interpolation GSL defines (gsl_interp2d_bicubic), restated in
tamcmc-c_b200/csrc/alm_grid.cpp.

Pins, strongest first:
  * tamcmc_alm_grids_make regenerates the reference's 30 shipped 1-degree grid files TOKEN FOR TOKEN (so tamcmc_host_alm
    reproduces the reference's Alm() integral to the 6 digits the files carry, on 121 500 nodes, and the file format is the
    reference's) -- needs /root/reference, i.e. the build container;
  * against the reference's OWN sources (activity.cpp, Alm_interpol.cpp, bilinear_interpol.cpp compiled where they lie,
    oracle/_ref): Alm() vs tamcmc_host_alm at 1e-13, Alm_interp_iter_preinitialised vs tamcmc_alm_grids_eval at 1e-13.  Boost
    and GSL are absent: the shim's stand-ins say so (oracle/eigen_shim/{boost,gsl}); the bicubic VALUE is therefore pinned on
    GSL's published algorithm written twice (term-by-term there, Hermite basis in the product, scipy's natural splines
    below), not on a GSL binary;
  * the reference's own test bar for the grids: interpolation vs direct integral within 1.5e-2
    (external/Alm/Alm_cpp/tests/unit_tests.cpp:443)."""
import gzip
import os

import numpy as np
import pytest

import _refshim

REF_GRIDS = "/root/reference/external/Alm/data/Alm_grids_CPP/1deg_grids"
needs_ref = pytest.mark.skipif(not (_refshim.available() and hasattr(_refshim.get().L, "ref_Alm")), reason="reference Alm sources not compiled here")
needs_ref_grids = pytest.mark.skipif(not os.path.isdir(REF_GRIDS), reason="the reference tree (shipped grids) is not present")


@pytest.fixture(scope="module")
def own_grid_dir(pkg, tmp_path_factory):
    """Grids written by the product's GridMaker: the GPU box has no /root/reference."""
    d = str(tmp_path_factory.mktemp("alm_grids"))
    pkg.AlmGrids.make(d, 0)
    pkg.AlmGrids.make(d, 2)
    return d


def _scipy_bicubic(x, y, z, xq, yq):
    """Independent restatement: natural cubic splines (scipy) for zx, zy, zxy at the nodes, then the bicubic Hermite patch."""
    from scipy.interpolate import CubicSpline
    zx = np.stack([CubicSpline(x, z[j], bc_type="natural")(x, 1) for j in range(len(y))])
    zy = np.stack([CubicSpline(y, z[:, i], bc_type="natural")(y, 1) for i in range(len(x))], axis=1)
    zxy = np.stack([CubicSpline(x, zy[j], bc_type="natural")(x, 1) for j in range(len(y))])
    i = min(np.searchsorted(x, xq, side="right") - 1, len(x) - 2)
    j = min(np.searchsorted(y, yq, side="right") - 1, len(y) - 2)
    dx, dy = x[i + 1] - x[i], y[j + 1] - y[j]
    t, u = (xq - x[i]) / dx, (yq - y[j]) / dy
    h = lambda s: np.array([2 * s**3 - 3 * s**2 + 1, -2 * s**3 + 3 * s**2, s**3 - 2 * s**2 + s, s**3 - s**2])
    ht, hu = h(t), h(u)
    r = 0.0
    for b in range(2):
        for a in range(2):
            r += (z[j + b, i + a] * ht[a] * hu[b] + zx[j + b, i + a] * dx * ht[2 + a] * hu[b]
                  + zy[j + b, i + a] * dy * ht[a] * hu[2 + b] + zxy[j + b, i + a] * dx * dy * ht[2 + a] * hu[2 + b])
    return r


def test_bicubic_against_independent_restatement(pkg, own_grid_dir):
    G = pkg.AlmGrids(own_grid_dir)
    rng = np.random.default_rng(5)
    for (l, m, fc) in [(1, 0, 0), (1, 1, 2), (2, 1, 0), (2, 2, 2), (3, 0, 0), (3, 3, 2)]:
        x, y, z = G.nodes(l, m, fc)
        assert x.shape == (90,) and y.shape == (45,) and z.shape == (45, 90)
        assert np.array_equal(G.nodes(l, -m, fc)[2], z)                         # m and -m share a grid (Alm_interpol.cpp:204)
        for i in (0, 7, 44, 89):
            for j in (0, 3, 44):
                assert G(l, m, x[i], y[j], fc) == pytest.approx(z[j, i], abs=1e-15)   # interpolation: exact at the nodes
        for _ in range(40):
            xq, yq = rng.uniform(x[0], x[-1]), rng.uniform(y[0], y[-1])
            assert G(l, m, xq, yq, fc) == pytest.approx(_scipy_bicubic(x, y, z, xq, yq), rel=1e-11, abs=1e-13)
    # outside the grid GSL raises GSL_EDOM (the reference's process aborts): NaN; bad (l, m): -9998 (Alm_interpol.cpp:194-202)
    assert np.isnan(G(1, 0, 1.7, 0.2, 0)) and np.isnan(G(1, 0, 0.5, 0.9, 0)) and np.isnan(G(1, 0, 0.5, 0.2, 1))
    assert G(0, 0, 0.5, 0.2, 0) == -9998 and G(4, 0, 0.5, 0.2, 0) == -9998 and G(2, 3, 0.5, 0.2, 0) == -9998


def test_grid_interpolation_vs_direct_integral_at_the_reference_tolerance(pkg, own_grid_dir):
    """external/Alm/Alm_cpp/tests/unit_tests.cpp:443: |interpolated - integrated| <= 1.5e-2 on random (theta0, delta)."""
    G = pkg.AlmGrids(own_grid_dir)
    rng = np.random.default_rng(6)
    worst = 0.0
    for _ in range(3000):
        l = int(rng.integers(1, 4)); m = int(rng.integers(-l, l + 1)); fc = int(rng.choice([0, 2]))
        t0, de = rng.uniform(0.0, np.pi / 2), rng.uniform(0.0, np.pi / 4)
        worst = max(worst, abs(G(l, m, t0, de, fc) - pkg.host_alm(l, m, t0, de, fc)))
    assert worst < 1.5e-2
    assert worst < 6e-3                     # what the grids' Readme promises for the 1-degree resolution


def test_expander_with_grids_needs_no_callback(pkg, oracle, own_grid_dir):
    """C3 from a reference parameter vector: tamcmc_host_expand_ajAlm with tamcmc_alm_grids_eval as the Alm provider gives the
    row the oracle's model 21 implies when it is fed the same grid values."""
    G = pkg.AlmGrids(own_grid_dir)
    for decompose, fc in [(-1, 0), (0, 2), (1, 0), (2, 2)]:
        rng = np.random.default_rng(40 + decompose)
        params, pl = pkg.synth.ajalm_params(rng, Nmax=6, lmax=3, f0=1000.0, dnu=70.0, decompose_Alm=decompose, filter_code=fc, trunc_c=20.0,
                                            theta0=rng.uniform(20, 80), delta=rng.uniform(2, 40))
        x = pkg.synth.freq_axis(20000, 950.0, 0.02)
        rc, M, tr = oracle.call_model(21, params, pl, x, alm=lambda l, m, t0, de, f, user: G(l, m, t0, de, f), trace=True)
        assert rc == 0
        row, nm = pkg.expand_ajAlm(params, pl, int(pl[2:6].sum()), alm=G)
        rc, M2, tr2 = oracle.mode_table_model(row, int(pl[8]), 0, x, trace=True)
        assert rc == 0 and all(np.array_equal(a, b) for a, b in zip(tr, tr2))
        assert np.max(np.abs(M2 - M) / np.abs(M)) < 1e-13
        row_int, _ = pkg.expand_ajAlm(params, pl, int(pl[2:6].sum()))             # direct integral instead of the grid
        assert not np.array_equal(row, row_int)                                    # ... is a (slightly) different model


@needs_ref_grids
def test_gridmaker_regenerates_the_shipped_grids_token_for_token(pkg, own_grid_dir):
    ntok = 0
    for f in ("gate", "triangle"):
        names = sorted(os.listdir(os.path.join(REF_GRIDS, f)))
        assert names == sorted(os.listdir(os.path.join(own_grid_dir, f))) and len(names) == 15
        for fn in names:
            a = gzip.open(os.path.join(own_grid_dir, f, fn), "rt").read().split()
            b = gzip.open(os.path.join(REF_GRIDS, f, fn), "rt").read().split()
            assert a == b, (f, fn)
            ntok += len(a)
    assert ntok == 30 * (90 + 45 + 1 + 45 * 90)


@needs_ref
def test_host_alm_against_the_reference_integral(pkg):
    """tamcmc_host_alm vs the reference's own Alm() (activity.cpp:221-246, GaussLegendre2D.hpp order 64)."""
    R = _refshim.get()
    rng = np.random.default_rng(7)
    for _ in range(400):
        l = int(rng.integers(1, 4)); m = int(rng.integers(-l, l + 1)); fc = int(rng.choice([0, 2]))
        t0, de = rng.uniform(0.0, np.pi / 2), rng.uniform(1e-3, np.pi / 4)
        a, b = pkg.host_alm(l, m, t0, de, fc), R.Alm(l, m, t0, de, fc)
        assert a == pytest.approx(b, rel=1e-12, abs=1e-15), (l, m, t0, de, fc)
    assert R.Alm(2, 1, 0.7, 0.0, 0) == 0.0 == pkg.host_alm(2, 1, 0.7, 0.0, 0)
    assert R.Alm(2, 3, 0.7, 0.1, 0) == -10.0 == pkg.host_alm(2, 3, 0.7, 0.1, 0)      # |m| > l (activity.cpp:240-243)


@needs_ref
@needs_ref_grids
def test_grid_reader_and_interpolator_against_the_reference_sources(pkg):
    """loadAllData + flatten_grid + init_2dgrid + Alm_interp_iter_preinitialised of the reference (its own .cpp files) on the
    shipped grids vs tamcmc_alm_grids_load / _eval."""
    R = _refshim.get()
    assert R.alm_grids_load(REF_GRIDS) == 0
    G = pkg.AlmGrids(REF_GRIDS)
    rng = np.random.default_rng(8)
    for fc in (0, 2):
        for l in (1, 2, 3):
            for m in range(-l, l + 1):
                x, y, z = G.nodes(l, m, fc)
                for i, j in [(0, 0), (89, 44), (13, 7), (50, 30)]:
                    assert R.Alm_interp(l, m, x[i], y[j], fc) == pytest.approx(z[j, i], abs=2e-14)      # same file, same cell (rounding of the 16-term sum)
                for _ in range(60):
                    t0, de = rng.uniform(0.0, x[-1]), rng.uniform(0.0, y[-1])
                    assert G(l, m, t0, de, fc) == pytest.approx(R.Alm_interp(l, m, t0, de, fc), rel=1e-12, abs=1e-14)
