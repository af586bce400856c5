#!/usr/bin/env python
"""C2 step time under three cache regimes (one GPU): (a) L2 flushed before every timed step (what bench.py reports), (b) nothing
between steps (an MCMC run's steady state: 6 MB of inputs and the kernels' code stay in the 126 MB L2), (c) K independent C2 stars
evaluated round-robin, K x 6 MB > L2 (inputs cold like (a), code and small tables warm like (b)).  Prints one JSON line."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stars", type=int, default=48)
    ap.add_argument("--steps", type=int, default=480)
    a = ap.parse_args()
    import torch
    import __graft_entry__ as g
    pkg = g.load_package()
    synth = pkg.synth
    T = synth.tcoefs(bench.NCHAINS, bench.LAMBDA_T)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctxs, dps, dLs = [], [], []
    for s in range(a.stars):
        rng, params, pl, x = bench.make_star(synth, s)
        if s == 0:
            with pkg.Context(pkg.Star(3, pl, len(params), x, np.ones_like(x)), 1, [1.0]) as c0:
                M = c0.model(params)
        y = synth.chi2_2dof_spectrum(rng, M)              # (one model shape for all: only the cache behaviour matters here)
        ctx = pkg.Context(pkg.Star(3, pl, len(params), x, y), bench.NCHAINS, T)
        P = ctx.pack_params([synth.perturb_chains(rng, params, pl, bench.NCHAINS)])
        ctxs.append(ctx)
        dps.append(torch.tensor(P, device="cuda"))
        dLs.append(torch.zeros(bench.NCHAINS, dtype=torch.float64, device="cuda"))
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")

    def run(mode):
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
        for i in range(a.stars if mode == "rotate" else 5):
            k = i % a.stars if mode == "rotate" else 0
            ctxs[k].eval_device(dps[k].data_ptr(), dLs[k].data_ptr(), stream=stream.cuda_stream)
        torch.cuda.synchronize()
        for i in range(a.steps):
            k = i % a.stars if mode == "rotate" else 0
            if mode == "flush":
                flush.zero_()
            ev0[i].record(stream)
            ctxs[k].eval_device(dps[k].data_ptr(), dLs[k].data_ptr(), stream=stream.cuda_stream)
            ev1[i].record(stream)
        torch.cuda.synchronize()
        t = np.array([x.elapsed_time(y) for x, y in zip(ev0, ev1)]) * 1e3
        return {"us_per_step_mean": float(t.mean()), "us_median": float(np.median(t)), "evals_per_s": bench.NCHAINS / (t.mean() * 1e-6)}

    out = {"workload": bench.WORKLOAD, "stars_in_rotation": a.stars, "input_bytes_in_rotation": a.stars * 24 * bench.NBINS,
           "flush": run("flush"), "warm": run("warm"), "rotate": run("rotate")}
    print(json.dumps(out))
    for c in ctxs:
        c.close()


if __name__ == "__main__":
    main()
