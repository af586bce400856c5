#!/usr/bin/env python
"""Turns the scratch outputs of profiles/gpu_round.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/<round>/:
bench lines, launch list + per-kernel shares of a step, the `ncu --set full` summary of the fused kernel, the DRAM traffic
bench.py quotes in `roofline.traffic` (profiles/roofline_traffic.json), the per-CTA timeline.
usage: python profiles/summarise_round.py <tag> <round-dir>      e.g.  r1c r1"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rnd = sys.argv[1], sys.argv[2]
src = os.path.join(ROOT, "gpurun_out")
dst = os.path.join(ROOT, "profiles", rnd)
os.makedirs(dst, exist_ok=True)
for a, b in (("bench.json", "bench_default.json"), ("bench_reference.json", "bench_reference.json"), ("configs.jsonl", "bench_configs.jsonl"),
             ("launches.csv", "launches_bench_steps5.csv"), ("trace.txt", "trace_timeline.txt"), ("l2_modes.json", "l2_modes.json"),
             ("rgb_launches.csv", "launches_rgb_c4.csv")):
    p = os.path.join(src, "%s_%s" % (tag, a))
    if os.path.exists(p):
        shutil.copy(p, os.path.join(dst, b))

# ---- launch list -> shares of one step (cold-cache, serialised: compare shares, not absolutes) ----
p = os.path.join(src, tag + "_launches.csv")
if os.path.exists(p):
    rows = list(csv.reader(open(p)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, body = rows[h], rows[h + 1:]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in body:
        if len(r) > iv:
            k = r[ik].split("(")[0].split("::")[-1][-48:]
            agg.setdefault(k, [0, 0.0])
            agg[k][0] += 1
            agg[k][1] += float(r[iv])
    step = {k: v for k, v in agg.items() if "expand" in k or "ksi_kernel" in k or ("whittle_kernel<0" in k) or "whittle_kernel<(bool)0" in k}
    tot = sum(v[1] / v[0] for v in step.values()) or 1.0
    with open(os.path.join(dst, "launch_shares.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none, command: python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra\n")
        f.write("per-launch times are cold-cache and serialised; one evaluation (step) = tamcmc_expand_kernel + tamcmc_whittle_kernel<0, bins per thread>\n\n")
        for k, v in agg.items():
            f.write("%-50s launches %4d  avg %9.2f us\n" % (k, v[0], v[1] / v[0] / 1e3))
        f.write("\nshare of one step:\n")
        for k, v in step.items():
            f.write("  %-48s %5.1f %%\n" % (k, 100 * (v[1] / v[0]) / tot))
    print(open(os.path.join(dst, "launch_shares.txt")).read())

# ---- the red-giant set-up kernels (tamcmc_gpu_rgb_expand, C4): launch list and the full capture of the search / pairs kernels ----
p = os.path.join(src, tag + "_rgb_launches.csv")
if os.path.exists(p):
    rows = list(csv.reader(open(p)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, body = rows[h], rows[h + 1:]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in body:
        if len(r) > iv:
            k = r[ik].split("(")[0].split("::")[-1]
            agg.setdefault(k, [0, 0.0])
            agg[k][0] += 1
            agg[k][1] += float(r[iv])
    with open(os.path.join(dst, "launch_shares_rgb_c4.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tamcmc_rgb, command: python profiles/bench_configs.py --configs c4 --steps 4\n")
        f.write("one tamcmc_gpu_rgb_expand call (10 chains) = ksi_max + ksi_top on one stream, search + pairs + compact on the other\n\n")
        for k, v in agg.items():
            f.write("%-40s launches %4d  avg %9.2f us\n" % (k, v[0], v[1] / v[0] / 1e3))
    print(open(os.path.join(dst, "launch_shares_rgb_c4.txt")).read())
rep = os.path.join(src, tag + "_rgb.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "ncu_summary.py"), rep], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    open(os.path.join(dst, "rgb_ncu_full_summary.txt"), "w").write(out)

# ---- full capture of the expander ----
rep = os.path.join(src, tag + "_expand.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "ncu_summary.py"), rep], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    open(os.path.join(dst, "expand_ncu_full_summary.txt"), "w").write(out)

# ---- full capture of the fused kernel ----
rep = os.path.join(src, tag + "_whittle.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "ncu_summary.py"), rep], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    open(os.path.join(dst, "whittle_ncu_full_summary.txt"), "w").write(out)
    raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout)))
    hdr = raw[0]
    def col(name, row):
        return float(row[hdr.index(name)].replace(",", "")) if name in hdr else None
    def unit(name):
        return raw[1][hdr.index(name)] if name in hdr else ""
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = [col("dram__bytes_read.sum", r) * scale.get(unit("dram__bytes_read.sum"), 1.0) for r in raw[2:]]
    wr = [col("dram__bytes_write.sum", r) * scale.get(unit("dram__bytes_write.sum"), 1.0) for r in raw[2:]]
    fp64 = [col("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", r) for r in raw[2:]]
    info = {"source": "profiles/%s/whittle_ncu_full_summary.txt (ncu --set full, %d launches of tamcmc_whittle_kernel, C2)" % (rnd, len(rd)),
            "whittle_dram_bytes_read_per_launch": sum(rd) / len(rd), "whittle_dram_bytes_write_per_launch": sum(wr) / len(wr),
            "whittle_dram_bytes_per_launch": (sum(rd) + sum(wr)) / len(rd),
            "whittle_fp64_pipe_active_pct_of_elapsed": sum(fp64) / len(fp64) if fp64[0] is not None else None}
    json.dump(info, open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w"), indent=1)
    print(json.dumps(info, indent=1))
