"""The far-field folding of the fused kernel (tamcmc-c_b200/csrc/whittle.cu: far_series, producer_loop) restated in numpy:
the Taylor coefficients of a scaled Lorentzian 1 / ((s u + c)^2 + a) about the tile centre follow the Chebyshev-U recurrence
f_0 = h, f_1 = p h, f_{k+1} = p f_k - q f_{k-1} (h = 1/(c^2 + a), p = -2 s c h, q = s^2 h), and 20 terms at an expansion ratio
of 5 reproduce the component to the bound DESIGN.md states (<= 5e-13 of its own value).  CPU-only: this checks the algorithm
the kernel implements and the defaults in tamcmc_dev.h, not the kernel (tests/test_gpu_parity.py does that)."""
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEV_H = os.path.join(HERE, "..", "tamcmc-c_b200", "csrc", "tamcmc_dev.h")


def _defaults():
    src = open(DEV_H).read()
    nfar = int(re.search(r"#define TAMCMC_FAR_TERMS (\d+)", src).group(1))
    ratio = float(re.search(r"#define TAMCMC_FAR_RATIO_DEFAULT ([0-9.]+)", src).group(1))
    return nfar, ratio


def far_series(s, c, a, nterms, Q=(1.0, 0.0, 0.0)):
    """Coefficients of q(u) / ((s u + c)^2 + a), q(u) = Q0 + Q1 u + Q2 u^2 (the asymmetry factor of build_lorentzian.cpp:153-157),
    in the order the kernel forms them."""
    h = 1.0 / (c * c + a)
    p, q = -2.0 * s * c * h, s * s * h
    f = np.empty(nterms)
    f[0], f[1] = h, p * h
    for k in range(2, nterms):
        f[k] = p * f[k - 1] - q * f[k - 2]
    g = Q[0] * f
    g[1:] += Q[1] * f[:-1]
    g[2:] += Q[2] * f[:-2]
    return g


def horner(coef, u):
    acc = np.full_like(u, coef[-1])
    for ck in coef[-2::-1]:
        acc = acc * u + ck
    return acc


def test_defaults_match_the_documented_bound():
    nfar, ratio = _defaults()
    bound = (nfar + 1) * ratio ** (-nfar) * ((1 + 1 / ratio) / (1 - 1 / ratio)) ** 2
    assert bound < 1e-12          # 5e-13; two orders inside the 1e-10 parity bar for a folded component's OWN value


def test_recurrence_reproduces_far_lorentzians():
    nfar, ratio = _defaults()
    rng = np.random.default_rng(3)
    umax = 6.09                                             # half a 1536-bin tile of a 4-year Kepler spectrum, microHz
    u = np.linspace(-umax, umax, 1537)
    worst = 0.0
    for _ in range(400):
        gamma = rng.uniform(0.05, 8.0)
        height = 10.0 ** rng.uniform(-3, 3)
        d = rng.choice([-1.0, 1.0]) * ratio * umax * rng.uniform(1.0, 30.0)     # nu - xc: at or beyond the far threshold
        s = 2.0 / (gamma * np.sqrt(height))                  # expand.cu: scaled FAST form
        a = 1.0 / height
        c = -d * s
        exact = 1.0 / ((s * u + c) ** 2 + a)                 # = height / (1 + 4 (x - nu)^2 / gamma^2)
        ref = height / (1.0 + 4.0 * (u - d) ** 2 / gamma ** 2)
        assert np.max(np.abs(exact - ref) / ref) < 1e-12
        approx = horner(far_series(s, c, a, nfar), u)
        worst = max(worst, np.max(np.abs(approx - exact) / exact))
    assert worst < 1e-12, worst


def test_asymmetric_profile_is_the_cauchy_product():
    nfar, ratio = _defaults()
    rng = np.random.default_rng(4)
    umax = 6.09
    u = np.linspace(-umax, umax, 513)
    worst = 0.0
    for _ in range(200):
        gamma = rng.uniform(0.1, 8.0)
        height = 10.0 ** rng.uniform(-2, 2)
        fc = rng.uniform(600.0, 3000.0)
        asym = rng.uniform(-100.0, 100.0)
        d = rng.choice([-1.0, 1.0]) * ratio * umax * rng.uniform(1.0, 10.0)     # fc - xc (one component at the mode centre)
        xc = fc - d
        s, a = 2.0 / (gamma * np.sqrt(height)), 1.0 / height
        c = -d * s
        qa, qb, qc = asym / fc, (1.0 - asym) + xc * (asym / fc), (0.5 * gamma * asym / fc) ** 2       # ModeRec.qa, qb0 + xc qa, qc
        x = xc + u
        exact = height / (1.0 + 4.0 * (x - fc) ** 2 / gamma ** 2) * ((1.0 + asym * (x / fc - 1.0)) ** 2 + (0.5 * gamma * asym / fc) ** 2)
        approx = horner(far_series(s, c, a, nfar, (qb * qb + qc, 2.0 * qa * qb, qa * qa)), u)
        worst = max(worst, np.max(np.abs(approx - exact) / np.max(np.abs(exact))))
    assert worst < 1e-11, worst
