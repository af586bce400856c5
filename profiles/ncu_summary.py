"""Summarise an .ncu-rep (raw metrics + per-source-line samples) into text. Usage: ncu_summary.py rep [kernel_substr]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
def run(args):
    return subprocess.run(["ncu", "-i", rep] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
raw = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
hdr = raw[0]
keys = ['gpu__time_duration.sum','sm__cycles_elapsed.max','sm__cycles_active.avg','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread',
        'launch__waves_per_multiprocessor','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__warps_eligible.avg.per_cycle_active','smsp__warps_active.avg.per_cycle_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active']
for row in raw[2:]:
    print("== kernel:", row[hdr.index('Kernel Name')][:80])
    for k in keys:
        if k in hdr: print("  %-70s %s %s" % (k, row[hdr.index(k)], raw[1][hdr.index(k)]))
    st = []
    for i, h in enumerate(hdr):
        if 'smsp__average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            try: st.append((float(row[i]), h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
            except: pass
    print("  stalls (warp-cycles per issue):", ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:9]))
mix = list(csv.reader(io.StringIO(run(["--page", "source", "--csv", "--print-source", "cuda,sass"]))))
hi = None
for i, r in enumerate(mix):
    if r and r[0] == 'Line No': hi = i; break
if hi is not None:
    h = mix[hi]; ismp = h.index('# Samples'); iex = h.index('Instructions Executed')
    agg = collections.OrderedDict(); cur = None
    for r in mix[hi+1:]:
        if len(r) <= iex: continue
        if r[0].isdigit(): cur = (int(r[0]), r[1][:100]); agg.setdefault(cur, [0, 0]); continue
        if cur and r[ismp].isdigit(): agg[cur][0] += int(r[ismp]); agg[cur][1] += int(r[iex])
    tot = sum(v[0] for v in agg.values()) or 1; totx = sum(v[1] for v in agg.values()) or 1
    print("== per source line (samples %, executed warp-instr %)  total samples", tot, "executed", totx)
    for (ln, src), (s, e) in sorted(agg.items()):
        if s > tot * 0.006 or e > totx * 0.006: print("%4d smp %5.1f%% exe %5.1f%%  %s" % (ln, 100*s/tot, 100*e/totx, src))
