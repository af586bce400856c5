"""Seeded fuzz of the GPU path against the oracle: random spectrum sizes, ladders, widths from sub-bin to very broad,
truncation coefficients, signs, zero heights, asymmetries, inclinations at the ends of the range -- per model id.  Bin
windows bit-exact, model and logL to 1e-10, window errors flagged on the same chains as the oracle's."""
import numpy as np
import pytest

import _cases

RTOL = 1e-10


def _fuzz_case(synth, model_id, seed):
    rng = np.random.default_rng(7000 + 31 * seed + model_id)
    N = int(rng.choice([3, 17, 700, 1536, 1537, 4000, 9215, 30001, 70000]))
    if model_id in (3, 6, 7, 8, 12, 13, 23):
        N = max(N, 700)
    step = float(rng.choice([synth.RESOL_4YR, 4 * synth.RESOL_4YR, 0.03, 0.25]))
    Nmax = int(rng.integers(2, 12))
    lmax = int(rng.integers(0, 4)) if model_id not in (6,) else int(rng.integers(2, 4))
    wmin = float(rng.choice([1e-3, 0.05, 0.5]))
    wmax = wmin * float(rng.choice([2.0, 20.0, 400.0]))
    kw = dict(N=N, x0=float(rng.choice([0.5, 40.0, 900.0, 4000.0])), step=step, Nmax=Nmax, lmax=lmax,
              asym=float(rng.choice([0.0, 0.0, 25.0, -80.0])), do_amp=int(rng.integers(0, 2)),
              inc=float(rng.choice([0.0, 90.0, 1e-9, rng.uniform(0, 90)])), a1=float(rng.choice([0.0, 0.4, 3.0, 25.0])),
              trunc_c=float(rng.choice([0.5, 5.0, 30.0, 1e4])), wmin=wmin, wmax=wmax)
    params, pl, x = _cases.ms_case(synth, model_id, seed, **kw)
    if model_id in (3, 6, 7, 8, 12, 23) and rng.random() < 0.5:
        params = params.copy()
        k = rng.integers(0, int(pl[0]), 2)
        params[k[0]] = 0.0                        # a dead radial order
        params[k[1]] = -params[k[1]]              # heights enter through std::abs
    return params, pl, x, rng


@pytest.mark.gpu
@pytest.mark.parametrize("model_id", _cases.ALL_MODELS)
def test_fuzz_against_oracle(pkg, oracle, model_id):
    checked = 0
    for seed in range(14):
        params, pl, x, rng = _fuzz_case(pkg.synth, model_id, seed)
        P = pkg.synth.perturb_chains(rng, params, pl, 3)
        P[2, -2] = params[-2] * float(rng.choice([1.0, 0.3, 2.5]))           # another truncation coefficient on one chain
        T = pkg.synth.tcoefs(3, 1.7)
        Ms, rcs, trs = [], [], []
        for r in P:
            rc, M, tr = oracle.call_model(model_id, r, pl, x, trace=True)
            Ms.append(M); rcs.append(rc); trs.append(tr)
        good = [i for i in range(3) if rcs[i] == 0 and np.all(np.isfinite(Ms[i])) and np.all(Ms[i] > 0)]
        y = (Ms[good[0]] if good else np.ones_like(x)) * rng.exponential(1.0, len(x))
        _, L_ref = oracle.eval_chains(model_id, P, pl, x, y, T)
        with pkg.Context(pkg.Star(model_id, pl, len(params), x, y), 3, T) as ctx:
            L, st = ctx.eval(P, raise_on_error=False)
            for i in range(3):
                if rcs[i] != 0:
                    assert st[0, i] & pkg.CHAIN_WINDOW, (model_id, seed, i)      # the reference would have exit()ed here
                    assert np.isnan(L[0, i])
                    continue
                if i not in good:
                    continue
                assert st[0, i] == 0, (model_id, seed, i, st)
                rcw, wl, w0, w1 = ctx.windows(P[i])
                assert np.array_equal(wl, trs[i][0]) and np.array_equal(w0, trs[i][1]) and np.array_equal(w1, trs[i][2]), (model_id, seed, i)
                assert np.max(np.abs(ctx.model(P[i]) - Ms[i]) / np.abs(Ms[i])) < RTOL, (model_id, seed, i)
                assert abs(L[0, i] - L_ref[i]) <= RTOL * abs(L_ref[i]), (model_id, seed, i, L[0, i], L_ref[i])
                checked += 1
    assert checked >= 20
