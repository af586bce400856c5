"""The exchange step of a bin-sharded spectrum (BASELINE config C3) behind the C ABI: every rank's fused kernel
writes its chains' local sums into the peers' exchange buffers (CUDA IPC peer memory), waits for their flags, and finalises the
log-likelihood of the WHOLE spectrum on the device -- no NCCL launch, no host round trip.

Two ranks = two PROCESSES (like torchrun's one process per GPU).  On a one-GPU box both use device 0 (the peer mapping is then a
same-device IPC mapping and the two persistent kernels time-slice), on a multi-GPU box rank r uses GPU r."""
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
    dev = rank % torch.cuda.device_count()
    params, pl, x = _cases.ms_case(pkg.synth, 3, seed=11, N=30000, asym=9.0)
    rc, M, tr = O.call_model(3, params, pl, x, trace=True)
    rng = np.random.default_rng(3)
    y = pkg.synth.chi2_2dof_spectrum(rng, M)
    Nch = 4
    T = pkg.synth.tcoefs(Nch, 1.7)
    lo, hi = shard.bin_shards(len(x), world, shard.bin_work(len(x), *tr))[rank]
    ok = True
    with pkg.Context(pkg.Star.shard(3, pl, len(params), x, y, lo, hi), Nch, T, device=dev) as ctx:
        mine = torch.from_numpy(ctx.exchange_handle().copy())
        allh = [torch.zeros(64, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(allh, mine)
        ctx.exchange_attach(rank, world, torch.stack(allh).numpy())
        dist.barrier()
        results = []
        for step in range(4):
            P = pkg.synth.perturb_chains(np.random.default_rng(100 + step), params, pl, Nch)     # same proposals on every rank
            active = None if step != 2 else np.array([1, 0, 1, 1], dtype=np.uint8)                # a masked chain (prior = -inf)
            L, st = ctx.eval(P, active=active)
            rc, L_ref = O.eval_chains(3, P, pl, x, y, T)
            live = np.ones(Nch, bool) if active is None else active.astype(bool)
            ok = ok and bool(np.max(np.abs(L[0][live] - L_ref[live]) / np.abs(L_ref[live])) < 1e-10)
            ok = ok and bool(np.all(np.isnan(L[0][~live]))) and bool(np.all(st[0][live] == 0))
            results.append(L[0].copy())
        # the ranks must hold the SAME bits (sums taken in rank order on every rank)
        mineL = torch.from_numpy(np.nan_to_num(np.stack(results), nan=-1.0))
        allL = [torch.zeros_like(mineL) for _ in range(world)]
        dist.all_gather(allL, mineL)
        ok = ok and all(torch.equal(allL[0], a) for a in allL)
        dist.barrier()
    with open(os.path.join(out_dir, "rank%d.txt" % rank), "w") as f:
        f.write("%d %d %d\n" % (ok, lo, hi))
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_rank_exchange_through_peer_memory(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rows = [open(tmp_path / ("rank%d.txt" % r)).read().split() for r in range(world)]
    assert all(r[0] == "1" for r in rows), rows
    assert int(rows[0][1]) == 0 and rows[0][2] == rows[1][1] and int(rows[1][2]) == 30000
