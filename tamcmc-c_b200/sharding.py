"""Host-side partitioning for multi-GPU runs (one process per GPU).

* independent stars / slices / chain ensembles: star s -> rank s mod G, no communication
  (the reference already runs one process per star: scripts/slurm/job.sh);
* one large spectrum: contiguous bin ranges per rank, balanced by the prefix sum of the per-bin
  component count (not by bin count), every rank gets all mode tables; one FP64 sum-allreduce of
  Nchains partial sums S = sum(ln M + y/M) per step, then logL = -p*S/T (model_def.cpp:399-401).
"""
import numpy as np

TILE = 1536  # bins per tile of the fused kernel (TAMCMC_TILE, csrc/tamcmc_dev.h); shard boundaries are tile-aligned


def star_shard(nstars, rank, world):
    """Indices of the stars owned by `rank`."""
    return [s for s in range(nstars) if s % world == rank]


def bin_work(N, l, i0, i1, base=1.0):
    """Per-bin work estimate: `base` (background + Whittle terms) plus (2l+1) per covering mode."""
    d = np.zeros(N + 1)
    for ll, a, b in zip(l, i0, i1):
        d[a] += 2 * ll + 1
        d[b] -= 2 * ll + 1
    return base + np.cumsum(d[:N])


def bin_shards(N, world, work=None, align=TILE):
    """Contiguous [lo,hi) per rank with ~equal total work; boundaries multiples of `align`."""
    if work is None:
        work = np.ones(N)
    cs = np.concatenate([[0.0], np.cumsum(work)])
    bounds = [0]
    for r in range(1, world):
        target = cs[-1] * r / world
        b = int(np.searchsorted(cs, target))
        b = int(round(b / align) * align)
        b = min(max(b, bounds[-1]), N)
        bounds.append(b)
    bounds.append(N)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def finalize_logL(S, p, Tcoefs):
    """Tempered chi^2(2,2p) log-likelihood from the all-reduced sum S (likelihoods.cpp:23-25; p truncated
    to long and the division by Tcoefs[m] as model_def.cpp:399-401)."""
    return (-float(int(p)) * np.asarray(S, dtype=np.float64)) / np.asarray(Tcoefs, dtype=np.float64)
