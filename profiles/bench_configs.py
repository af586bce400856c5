#!/usr/bin/env python
"""Throughput of the OTHER BASELINE.json configs (bench.py measures C2, the config the metric is quoted on):

  C1  quick-start red-giant case: the reference's fixture 10722175 (slice 80-128 microHz, ~6100 bins), mixed-mode list
      resolved by the reference's own ARMM host code (tests/golden/reference_rgb_vectors.npz), 5 chains, lambda = 3.5
  C3  MS_Global ajAlm (gate filter, decompose_Alm = 1), 33 modes l <= 2, 10^6 bins, 10 chains; bin-sharded over the
      ranks when launched with torchrun (per-chain partial sums all-reduced with NCCL)
  C4  dense red-giant mode list: the 78-mode case of the same fixture, 10 chains
  C5  256 independent C2 stars x 10 chains, star-sharded over the ranks (no collective)

  python profiles/bench_configs.py [--configs c1,c3,c4,c5,env,driver] [--steps K] [--stars 256]
  python -m torch.distributed.run --nproc-per-node N profiles/bench_configs.py --configs c3,c5 ...

One JSON line per config on rank 0.  Device-resident timing (CUDA events on the launch stream, max over ranks) and the
end-to-end C-ABI call with host buffers; every config is first checked against the CPU oracle on a sub-sample."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timed(torch, ctx, P_host, steps, dist, extra=None):
    """-> (device ms/step, e2e ms/step) for one context; `extra(d_out)` runs on the stream after each device eval."""
    stream = torch.cuda.current_stream()
    d_params = torch.tensor(P_host, device="cuda")
    n = P_host.shape[0] * P_host.shape[1]
    d_out = torch.zeros(n, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    raw = extra is not None

    def step():
        ctx.eval_device(d_params.data_ptr(), d_out.data_ptr(), raw_sum=raw, stream=stream.cuda_stream)
        if extra:
            extra(d_out)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    for i in range(steps):
        flush.zero_()
        ev0[i].record(stream)
        step()
        ev1[i].record(stream)
    torch.cuda.synchronize()
    dev = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1)) / steps
    e2e = None
    if not raw:
        ctx.set_profiling(True)                      # CUDA event-record nodes between the kernels of the replayed graph (warm L2)
        for _ in range(10):
            ctx.eval(P_host)
        nprof, ems, wms = ctx.kernel_ms()
        ctx.set_profiling(False)
        timed.kernel_us = (1e3 * ems / max(nprof, 1), 1e3 * wms / max(nprof, 1))
        for _ in range(3):
            ctx.eval(P_host)
        t0 = time.perf_counter()
        for _ in range(steps):
            ctx.eval(P_host)
        e2e = (time.perf_counter() - t0) * 1e3 / steps
        # rows built in place in the context's pinned staging block (tamcmc_gpu_params_staging): the call skips its host-side copy
        S = ctx.params_staging()
        S[...] = P_host
        for _ in range(3):
            ctx.eval(S)
        t0 = time.perf_counter()
        for _ in range(steps):
            ctx.eval(S)
        timed.e2e_staged = (time.perf_counter() - t0) * 1e3 / steps
    if dist:
        t = torch.tensor([dev, e2e or 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev, e2e = float(t[0]), (float(t[1]) if e2e is not None else None)
    return dev, e2e, d_out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c3,c4,c5")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--stars", type=int, default=256)
    args = ap.parse_args()
    import torch
    import __graft_entry__ as g
    import _oracle
    pkg = g.load_package()
    synth = pkg.synth
    O = _oracle.get()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    torch.cuda.set_stream(torch.cuda.Stream())
    todo = args.configs.split(",")

    def emit(name, workload, evals_per_step, dev, e2e, pairs, extra=None):
        if rank != 0:
            return
        line = {"config": name, "workload": workload, "n_gpus": world, "evals_per_step": evals_per_step,
                "value": evals_per_step / (dev * 1e-3), "unit": "evals/s", "ms_per_step": dev,
                "e2e": {"value": evals_per_step / (e2e * 1e-3), "ms_per_step": e2e} if e2e else None,
                "pairs_per_step": pairs, "alg_tflops": (6.0 * pairs) / (dev * 1e-3) / 1e12 if pairs else None}
        if extra:
            line.update(extra)
        if getattr(timed, "kernel_us", None):
            line["expand_kernel_us"], line["fused_kernel_us"] = timed.kernel_us
            timed.kernel_us = None
        print(json.dumps(line), flush=True)

    # ------------------------------------------------------------------ C1 / C4: red giant, mode table
    if "c1" in todo or "c4" in todo:
        gold = np.load(os.path.join(ROOT, "tests", "golden", "reference_rgb_vectors.npz"))
        x, y = gold["x"], gold["y"]
        ncases = int(gold["ncases"])

        step_x = x[2] - x[1]
        cap = 110

        def params_of(i, rng):
            """Reference parameter vector of fixture case i (tests/golden/make_golden_rgb_from_reference_cpp.py), heights jittered
            by 2 % like the chains of a run."""
            params, pl = gold["params%d" % i].copy(), gold["plength%d" % i]
            params[:int(pl[0])] *= 1.0 + 0.02 * rng.standard_normal()
            return params, pl

        for name, picks, lam in (("c1", [0, 1, 2, 3, 0], 3.5), ("c4", [2] * 10, 1.7)):
            if name not in todo or rank != 0:
                continue
            rng = np.random.default_rng(1)
            vecs = [params_of(i, rng) for i in picks]
            # host expander: reference parameter vector -> mode-table row (ARMM mixed-mode solve, bias spline, zeta function)
            rows = np.stack([pkg.expand_rgb_v4(25, p, pl, step_x, cap)[0] for p, pl in vecs])
            nn = int(vecs[0][1][8])
            T = synth.tcoefs(len(picks), lam)
            rc, L_ref = O.mode_table_eval_chains(rows, nn, 1, x, y, T)
            star = pkg.Star(synth.MODEL_MODE_TABLE, synth.mode_table_plength(cap, nn, 1), rows.shape[1], x, y)
            with pkg.Context(star, len(picks), T, device=lr) as ctx:
                L, st = ctx.eval(rows)
                err = float(np.max(np.abs(L[0] - L_ref) / np.abs(L_ref)))
                assert (st == 0).all() and err < 1e-10, err
                pairs = ctx.pairs_last()
                dev, e2e, _ = timed(torch, ctx, ctx.pack_params([rows]), args.steps * 4, None)
                # the whole step a caller with reference parameter vectors pays: host solve of every chain + the batched evaluation
                nrep = max(args.steps // 10, 3)
                t0 = time.perf_counter()
                for _ in range(nrep):
                    rr = np.stack([pkg.expand_rgb_v4(25, p, pl, step_x, cap)[0] for p, pl in vecs])
                t_host = (time.perf_counter() - t0) * 1e3 / nrep
                t0 = time.perf_counter()
                for _ in range(nrep):
                    rr = np.stack([pkg.expand_rgb_v4(25, p, pl, step_x, cap)[0] for p, pl in vecs])
                    ctx.eval(rr)
                t_full = (time.perf_counter() - t0) * 1e3 / nrep
                # the same step with the pair loop and the zeta normalisation on the device (tamcmc_gpu_rgb_expand), rows written
                # straight into the context's staging block
                Pmat = np.stack([p for p, _ in vecs])
                stage = ctx.params_staging()[0]
                with pkg.RgbExpander(25, vecs[0][1], step_x, cap, len(picks), device=lr) as rx:
                    rows_d, nm_d, st_d, path_d = rx.expand(Pmat)
                    fc_h = rows[:, 4 + nn:].reshape(len(picks), cap, 20)[:, :, 1]
                    fc_d = rows_d[:, 4 + nn:].reshape(len(picks), cap, 20)[:, :, 1]
                    rgb_dev = {"status_ok": bool((st_d == 0).all()), "chains_on_device": int((path_d == 0).sum()),
                               "fc_identical_to_host_solver": int((fc_h == fc_d).sum()), "fc_total": int((fc_h != 0).sum()),
                               "fc_max_diff_ulp": float(np.max(np.abs(fc_h - fc_d) / np.spacing(np.maximum(np.abs(fc_h), 1e-300)))),
                               "rows_max_rel_diff": float(np.max(np.abs(rows - rows_d) / np.maximum(np.abs(rows), 1e-300)))}
                    L_d, _ = ctx.eval(rows_d)
                    rgb_dev["logL_max_rel_err_vs_oracle"] = float(np.max(np.abs(L_d[0] - L_ref) / np.abs(L_ref)))
                    nrep_d = max(args.steps, 20)
                    for _ in range(3):
                        rx.expand(Pmat, rows_out=stage)
                        ctx.eval(stage)
                    acc = np.zeros(4)
                    t0 = time.perf_counter()
                    for _ in range(nrep_d):
                        rx.expand(Pmat, rows_out=stage)
                        tt = rx.timings()
                        acc += [tt["prepare_ms"], tt["device_ms"], tt["finish_ms"], tt["total_ms"]]
                        ctx.eval(stage)
                    t_dev_full = (time.perf_counter() - t0) * 1e3 / nrep_d
                    rgb_dev.update({"ms_per_step": t_dev_full, "value": len(picks) / (t_dev_full * 1e-3),
                                    "expand_ms": {"prepare_host": acc[0] / nrep_d, "device": acc[1] / nrep_d, "finish_host": acc[2] / nrep_d,
                                                  "total": acc[3] / nrep_d}})
            emit(name, "red-giant fixture 10722175, %d bins, %d chains, %d-%d modes/chain, model 25 from reference parameter vectors"
                 % (len(x), len(picks), int(rows[:, 0].min()), int(rows[:, 0].max())), len(picks), dev, e2e, pairs,
                 {"max_rel_err_vs_oracle": err,
                  "e2e_rows_in_staging": {"ms_per_step": timed.e2e_staged, "value": len(picks) / (timed.e2e_staged * 1e-3)},
                  "host_solve_ms_per_step": t_host, "e2e_with_host_solve": {"ms_per_step": t_full, "value": len(picks) / (t_full * 1e-3),
                                                                            "host_share": t_host / t_full, "host_threads": os.cpu_count()},
                  "e2e_with_device_solve": rgb_dev,
                  "note": "single GPU (replicas only: SURVEY.md 8e). value / e2e: mode-table rows already resolved; e2e_with_host_solve: "
                          "tamcmc_host_expand_rgb_v4 (ARMM solver + zeta function, OpenMP) for every chain inside the timed region"})

    # ------------------------------------------------------------------ C3: ajAlm, 10^6 bins, bin-sharded
    if "c3" in todo:
        from importlib import import_module
        shard = import_module("tamcmc_c_b200.sharding")
        rng = np.random.default_rng(3427720)
        N, Nch = 1000000, 10
        params, pl = synth.ajalm_params(rng, Nmax=11, lmax=2, f0=2100.0, dnu=103.0, decompose_Alm=1, filter_code=0,
                                        epsilon=5e-3, theta0=50.0, delta=20.0, trunc_c=30.0)
        x = synth.freq_axis(N, 100.0)
        capm = int(pl[2:6].sum())
        nn = int(pl[8])
        mpl = synth.mode_table_plength(capm, nn, 0)
        row0, _ = pkg.expand_ajAlm(params, pl, capm)
        with pkg.Context(pkg.Star(synth.MODEL_MODE_TABLE, mpl, len(row0), x, np.ones(N)), 1, [1.0], device=lr) as c0:
            M = c0.model(row0)
            _, wl, w0, w1 = c0.windows(row0)
        y = synth.chi2_2dof_spectrum(rng, M)
        P = synth.perturb_chains(rng, params, pl, Nch)
        rows = np.stack([pkg.expand_ajAlm(P[c], pl, capm)[0] for c in range(Nch)])
        T = synth.tcoefs(Nch, 1.7)
        rc, Mo = O.mode_table_model(rows[1], nn, 0, x)              # oracle check of one chain at full size
        L1_ref = O.call_likelihood(y, Mo, 1.0, T[1])
        lo, hi = shard.bin_shards(N, world, shard.bin_work(N, wl, w0, w1))[rank]
        star = pkg.Star.shard(synth.MODEL_MODE_TABLE, mpl, rows.shape[1], x, y, lo, hi) if world > 1 else \
            pkg.Star(synth.MODEL_MODE_TABLE, mpl, rows.shape[1], x, y)
        with pkg.Context(star, Nch, T, device=lr) as ctx:
            if world > 1:
                def allreduce(d_S):
                    dist.all_reduce(d_S, op=dist.ReduceOp.SUM)      # 80 bytes: the path's one exchange step
                dev, e2e, d_S = timed(torch, ctx, ctx.pack_params([rows]), args.steps, dist, extra=allreduce)
                # d_S was all-reduced in place after the last evaluation
                L = shard.finalize_logL(d_S.cpu().numpy(), 1.0, T)
            else:
                dev, e2e, d_L = timed(torch, ctx, ctx.pack_params([rows]), args.steps, None)
                L = d_L.cpu().numpy()
            pairs = ctx.pairs_last()
        if dist:
            pt = torch.tensor([float(pairs)], dtype=torch.float64, device="cuda")
            dist.all_reduce(pt)
            pairs = int(pt[0])
        err = abs(L[1] - L1_ref) / abs(L1_ref)
        assert err < 1e-10, err
        emit("c3", "MS_Global ajAlm (gate, decompose_Alm=1) via host expander, 33 modes l<=2, 10^6 bins, 10 chains", Nch, dev, e2e, pairs,
             {"max_rel_err_vs_oracle": float(err), "sharding": "bins, balanced by per-bin work, tile-aligned; NCCL sum-allreduce of 10 partial sums" if world > 1 else "none",
              "note": "host expander (tamcmc_host_expand_ajAlm) outside the timed region"})

    # ------------------------------------------------------------------ C5: 256 stars x 10 chains, star-sharded
    if "c5" in todo:
        from importlib import import_module
        shard = import_module("tamcmc_c_b200.sharding")
        import bench
        mine = shard.star_shard(args.stars, rank, world)
        T = synth.tcoefs(10, 1.7)
        stars, Ps = [], []
        p0, pl0 = synth.classic_params(np.random.default_rng(0))
        with pkg.Context(pkg.Star(3, pl0, len(p0), synth.freq_axis(bench.NBINS, 500.0), np.ones(bench.NBINS)), 1, [1.0], device=lr) as c0:
            for s in mine:
                rng = np.random.default_rng(12345 + s)
                # Dnu ~ U(60, 95): 20 radial orders stay inside the 500-2480 microHz spectrum (SURVEY.md 8d says U(60,130),
                # which pushes the upper orders of a 20-order comb past the last bin)
                params, pl = synth.classic_params(rng, dnu=rng.uniform(60.0, 95.0), f0=rng.uniform(560.0, 620.0))
                x = synth.freq_axis(bench.NBINS, 500.0)
                M = c0.model(params)
                y = synth.chi2_2dof_spectrum(rng, M)
                stars.append(pkg.Star(3, pl, len(params), x, y))
                Ps.append(synth.perturb_chains(rng, params, pl, 10))
        rc, L_ref = O.eval_chains(3, Ps[0], stars[0].plength, stars[0].x, stars[0].y, T)
        with pkg.Context(stars, 10, T, device=lr) as ctx:
            P_host = ctx.pack_params(Ps)
            L, st = ctx.eval(P_host)
            err = float(np.max(np.abs(L[0] - L_ref) / np.abs(L_ref)))
            assert (st == 0).all() and err < 1e-10, err
            pairs = ctx.pairs_last()
            dev, e2e, _ = timed(torch, ctx, P_host, max(args.steps // 5, 5), dist)
        if dist:
            pt = torch.tensor([float(pairs)], dtype=torch.float64, device="cuda")
            dist.all_reduce(pt)
            pairs = int(pt[0])
        emit("c5", "%d independent C2-like stars (Dnu 60-95 microHz) x 10 chains, 250k bins each, star-sharded" % args.stars,
             args.stars * 10, dev, e2e, pairs, {"max_rel_err_vs_oracle": err, "stars_per_gpu": len(mine)})
    # ------------------------------------------------------------------ envelope models (ids 0, 1): no Lorentzians
    if "env" in todo and rank == 0:
        import bench
        N = bench.NBINS
        T = synth.tcoefs(10, 1.7)
        x = np.arange(N) * (283.2 / N)
        for mid, make, name in ((0, synth.kallinger_gaussian_params, "model_Kallinger2014_Gaussian"), (1, synth.harvey_gaussian_params, "model_Harvey_Gaussian")):
            rng = np.random.default_rng(7 + mid)
            rows = np.stack([make(rng, jitter=0.02) for _ in range(10)])
            if mid == 0:
                rows[:, [4, 9, 12]] = 4.0          # super-Lorentzian slopes fixed at 4 as in the reference's .model files
            rc, M0 = O.call_model(mid, rows[0], synth.ENVELOPE_PLENGTH, x)
            y = M0 * rng.exponential(1.0, N)
            t0 = time.perf_counter()
            rc, L_ref = O.eval_chains(mid, rows, synth.ENVELOPE_PLENGTH, x, y, T, nthreads=os.cpu_count())
            cpu_s = time.perf_counter() - t0
            with pkg.Context(pkg.Star(mid, synth.ENVELOPE_PLENGTH, rows.shape[1], x, y), 10, T, device=lr) as ctx:
                P_host = ctx.pack_params(rows)
                L, st = ctx.eval(P_host)
                err = float(np.max(np.abs(L[0] - L_ref) / np.abs(L_ref)))
                assert (st == 0).all() and err < 1e-10, err
                dev, e2e, _ = timed(torch, ctx, P_host, args.steps, None)
            emit("env%d" % mid, "%s, %d bins from 0 microHz, 10 chains" % (name, N), 10, dev, e2e, 0,
                 {"max_rel_err_vs_oracle": err, "cpu_port_evals_per_s": 10 / cpu_s, "cpu_threads": os.cpu_count()})
    # ------------------------------------------------------------------ MCMC steps/s with the C++ driver (C2)
    if "driver" in todo and rank == 0:
        import subprocess
        import tempfile
        import bench
        import test_host_cpp
        exe = test_host_cpp._build_driver()
        rng, params, pl, x = bench.make_star(synth, 0)
        with pkg.Context(pkg.Star(3, pl, len(params), x, np.ones_like(x)), 1, [1.0], device=lr) as c0:
            M = c0.model(params)
        y = synth.chi2_2dof_spectrum(rng, M)
        T = synth.tcoefs(10, 1.7)
        with tempfile.TemporaryDirectory() as td:
            f = os.path.join(td, "c2.bin")
            hdr = np.concatenate([[3, len(x), 10, len(params), 1.0], pl.astype(float)])
            with open(f, "wb") as fh:
                for a in (hdr, x, y, T, np.tile(params, (10, 1)).ravel()):
                    fh.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
            r = subprocess.run([exe, f, "4000", "bench"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            r2 = subprocess.run([exe, f, "600", "batch", "32"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        d = json.loads(line[-1]) if line else {"error": r.stdout[-400:]}
        d.update({"config": "driver", "workload": "C2 star, fixed-seed adaptive Metropolis + parallel tempering (host/mcmc_driver.hpp): proposals, priors, "
                  "accept/reject, learning and swaps on the host, one tamcmc_gpu_eval per step"})
        print(json.dumps(d), flush=True)
        line = [l for l in r2.stdout.splitlines() if l.startswith("{")]
        d = json.loads(line[-1]) if line else {"error": r2.stdout[-400:]}
        d.update({"config": "driver_batch", "workload": "32 independent C2 stars driven by BatchDriver: one tamcmc_gpu_eval of 320 chains per step, "
                  "host halves one star per OpenMP thread"})
        print(json.dumps(d), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
