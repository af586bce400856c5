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


def bin_work(N, l, i0, i1, base=15.0):
    """Per-bin work estimate: `base` for the fixed per-tile phases (tile switch, background polynomial, Whittle terms) plus
    (2l+1) per covering mode.  base = 15: measured on B200 (profiles/trace_tiles.py, C2): a tile costs ~1.86 us + 0.12 us per
    listed component, i.e. the fixed part weighs as much as ~15 components."""
    d = np.zeros(N + 1)
    for ll, a, b in zip(l, i0, i1):
        d[a] += 2 * ll + 1
        d[b] -= 2 * ll + 1
    return base + np.cumsum(d[:N])


def bin_shards(N, world, work=None, align=TILE):
    """Contiguous [lo,hi) per rank with ~equal total work; inner boundaries are multiples of `align` (the full tile: also a
    multiple of the half-size tile the library picks for small contexts, csrc/capi.cu).  Every rank gets at least `align` bins
    (tamcmc_gpu_create refuses fewer than 2): a spectrum too short for that raises instead of handing some rank an empty
    shard while the others enter the exchange."""
    if world < 1 or N < 2:
        raise ValueError("bin_shards: need world >= 1 and N >= 2")
    if world > 1 and N < world * align:
        raise ValueError("bin_shards: %d bins cannot give %d ranks at least %d bins each; use fewer ranks" % (N, world, align))
    if work is None:
        work = np.ones(N)
    cs = np.concatenate([[0.0], np.cumsum(work)])
    bounds = [0]
    for r in range(1, world):
        target = cs[-1] * r / world
        b = int(np.searchsorted(cs, target))
        b = int(round(b / align) * align)
        lo = bounds[-1] + align                          # this rank's shard is at least `align` bins ...
        hi = ((N - (world - r) * align) // align) * align  # ... and so is every later one
        b = min(max(b, lo), hi)
        bounds.append(b)
    bounds.append(N)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def finalize_logL(S, p, Tcoefs, likelihood_id=0):
    """Tempered log-likelihood from the all-reduced raw sum S of the shards (tamcmc_gpu_eval_device with raw_sum = 1).
    likelihood_id 0, chi^2(2,2p): -p S / T with S = sum(ln M + y/M), p truncated to long (likelihoods.cpp:23-25,
    model_def.cpp:399-401); likelihood_id 1, chi_square: -(S / 2) / T with S = sum((y - M)^2 / sigma^2) (likelihoods.cpp:36-37,
    model_def.cpp:405)."""
    S = np.asarray(S, dtype=np.float64)
    T = np.asarray(Tcoefs, dtype=np.float64)
    if likelihood_id == 1:
        return ((-S) / 2) / T
    if likelihood_id != 0:
        raise ValueError("unknown likelihood id %r" % (likelihood_id,))
    return (-float(int(p)) * S) / T
