"""Multi-process host logic of the N>1 paths, world_size 2 over gloo on CPU:
  * star sharding: star s -> rank s mod G, no collective on the data path, gather of logL only for output;
  * bin sharding of one spectrum: tile-aligned ranges balanced by per-bin work, every rank sums ITS bins, one FP64
    sum-allreduce of Nchains partial sums S, then logL = -p*S/T (model_def.cpp:399-401).
The per-rank arithmetic is done by the CPU oracle here (no GPU in this test); the same code path runs with
tamcmc_gpu_eval_device(raw_sum=1) + NCCL on the GPU box (tests/test_gpu_parity.py::test_bin_sharded_sums_add_up, bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as g
    import _cases
    import _oracle
    from importlib import import_module
    pkg = g.load_package()
    shard = import_module("tamcmc_c_b200.sharding")
    O = _oracle.get()

    # ---- bin sharding ----
    params, pl, x = _cases.ms_case(pkg.synth, 3, seed=11, N=20000)
    rc, M, tr = O.call_model(3, params, pl, x, trace=True)
    rng = np.random.default_rng(3)
    y = pkg.synth.chi2_2dof_spectrum(rng, M)
    Nch = 3
    P = pkg.synth.perturb_chains(rng, params, pl, Nch)
    T = pkg.synth.tcoefs(Nch, 1.7)
    work = shard.bin_work(len(x), *tr)
    lo, hi = shard.bin_shards(len(x), world, work)[rank]
    S = torch.zeros(Nch, dtype=torch.float64)
    for c in range(Nch):
        rc, Mc = O.call_model(3, P[c], pl, x)       # every rank expands all modes; it only SUMS its own bins
        S[c] = float(np.sum(np.log(Mc[lo:hi])) + np.sum(y[lo:hi] / Mc[lo:hi]))
    dist.all_reduce(S, op=dist.ReduceOp.SUM)
    L = shard.finalize_logL(S.numpy(), 1.0, T)
    rc, L_ref = O.eval_chains(3, P, pl, x, y, T)
    ok_bins = bool(np.max(np.abs(L - L_ref) / np.abs(L_ref)) < 1e-12)

    # ---- star sharding: 5 stars over 2 ranks, gather for output only ----
    nstars = 5
    mine = shard.star_shard(nstars, rank, world)
    local = torch.full((nstars,), float("nan"), dtype=torch.float64)
    for s in mine:
        ps, pls, xs = _cases.ms_case(pkg.synth, 3, seed=40 + s, N=4000, Nmax=3, lmax=2)
        rc, Ms = O.call_model(3, ps, pls, xs)
        local[s] = O.call_likelihood(Ms, Ms, 1.0, 1.0)
    gathered = [torch.zeros(nstars, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, local)
    full = torch.stack(gathered)
    owned = (~torch.isnan(full)).sum(0)
    ok_stars = bool((owned == 1).all()) and sorted(mine) == [s for s in range(nstars) if s % world == rank]
    with open(os.path.join(out_dir, "rank%d.txt" % rank), "w") as f:
        f.write("%d %d %d %d\n" % (ok_bins, ok_stars, lo, hi))
    dist.destroy_process_group()


def test_world2_bin_and_star_sharding(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rows = [open(tmp_path / ("rank%d.txt" % r)).read().split() for r in range(world)]
    assert all(r[0] == "1" and r[1] == "1" for r in rows), rows
    # contiguous cover of the spectrum, tile-aligned interior boundary
    assert int(rows[0][2]) == 0 and rows[0][3] == rows[1][2] and int(rows[1][3]) == 20000
    assert int(rows[0][3]) % 1536 == 0


def test_bin_shards_balance_work(pkg):
    from importlib import import_module
    shard = import_module("tamcmc_c_b200.sharding")
    N = 250000
    work = np.ones(N)
    work[60000:120000] += 80.0        # the mode region is ~80x heavier than the wings
    for world in (2, 4, 8):
        r = shard.bin_shards(N, world, work)
        assert r[0][0] == 0 and r[-1][1] == N and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        tot = [work[lo:hi].sum() for lo, hi in r]
        assert max(tot) / (sum(tot) / world) < 1.15          # within 15% of perfect balance
        assert all(lo % shard.TILE == 0 for lo, _ in r)


def test_bin_shards_never_hand_out_an_empty_shard(pkg):
    """Peaked work or a short spectrum: every rank still gets at least one full tile (tamcmc_gpu_create refuses N < 2), or the
    call raises before any rank enters the exchange; the chi_square finaliser (likelihoods.cpp:36-37, model_def.cpp:405)."""
    from importlib import import_module
    shard = import_module("tamcmc_c_b200.sharding")
    w = np.ones(100000)
    w[:2000] += 1e6                                    # all the work in the first two tiles
    r = shard.bin_shards(100000, 8, w)
    assert all(hi - lo >= shard.TILE for lo, hi in r) and r[0][0] == 0 and r[-1][1] == 100000
    assert all(a[1] == b[0] for a, b in zip(r, r[1:])) and all(lo % shard.TILE == 0 for lo, _ in r)
    assert shard.bin_shards(3 * shard.TILE + 5, 3) == [(0, 1536), (1536, 3072), (3072, 4613)]
    with pytest.raises(ValueError):
        shard.bin_shards(3000, 4)
    assert np.array_equal(shard.finalize_logL([2.0, 4.0], 1.0, [1.0, 2.0], likelihood_id=1), [-1.0, -1.0])
    assert np.array_equal(shard.finalize_logL([2.0, 4.0], 2.9, [1.0, 2.0]), [-4.0, -4.0])      # p truncated to long
