"""Per-CTA timeline of the fused kernel (needs a -DTAMCMC_TRACE build). Prints where CTAs spend their time."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g
pkg = g.load_package(); synth = pkg.synth
import bench
rng = np.random.default_rng(12345)
trunc = float(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1] != "c1" else 30.0
if len(sys.argv) > 1 and sys.argv[1] == "c1":
    # red-giant fixture (BASELINE config C1): mode table, 5 chains, ~6100 bins
    gold = np.load(os.path.join(ROOT, "tests", "golden", "reference_rgb_vectors.npz"))
    x, y = gold["x"], gold["y"]
    cap = max(len(gold["rows%d" % i]) for i in range(4)) + 2
    rows = []
    for i in (0, 1, 2, 3, 0):
        params, pl, r = gold["params%d" % i], gold["plength%d" % i], gold["rows%d" % i]
        o = int(pl[:8].sum()); nn = int(pl[8])
        rows.append(synth.mode_table_row(cap, abs(params[o + nn]), r[0, 13], r[0, 11], params[o:o + nn], r[:, :11]))
    rows = np.stack(rows)
    T = synth.tcoefs(5, 3.5)
    ctx = pkg.Context(pkg.Star(synth.MODEL_MODE_TABLE, synth.mode_table_plength(cap, nn, 1), rows.shape[1], x, y), 5, T)
    P = ctx.pack_params([rows])
else:
    params, pl = synth.classic_params(rng, trunc_c=trunc)
    if len(sys.argv) > 2 and int(sys.argv[2]): params[:20] = 0.0
    x = synth.freq_axis(bench.NBINS, 500.0)
    with pkg.Context(pkg.Star(3, pl, len(params), x, np.ones_like(x)), 1, [1.0]) as c0:
        M = c0.model(params)
    y = synth.chi2_2dof_spectrum(rng, M)
    T = synth.tcoefs(10, 1.7)
    ctx = pkg.Context(pkg.Star(3, pl, len(params), x, y), 10, T)
    P = ctx.pack_params([synth.perturb_chains(rng, params, pl, 10)])
for _ in range(5): ctx.eval(P)
n = 148
buf = np.zeros((n, 64), dtype=np.uint64)
ctx.eval(P)
rc = pkg.lib().tamcmc_gpu_debug_trace(ctx.h, buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), n)
assert rc == 0, rc
b = buf.astype(np.int64)
t0 = b[:, 0][b[:, 0] > 0].min()
start = b[:, 0] - t0; end = b[:, 1] - t0
print("CTAs with work:", (b[:, 0] > 0).sum())
print("kernel span (ns): first start 0, last start %d, first end %d, last end %d" % (start.max(), end[end > 0].min(), end.max()))
waits = []; ntiles = []; first_wait = []
for i in range(n):
    if b[i, 0] == 0: continue
    w = 0; k = 0
    for s in range(2, 64, 2):
        if b[i, s] == 0 or b[i, s + 1] == 0: break
        d = b[i, s + 1] - b[i, s]
        if k == 0: first_wait.append(d)
        w += d; k += 1
    waits.append(w); ntiles.append(k - 1)
waits = np.array(waits); ntiles = np.array(ntiles); first_wait = np.array(first_wait)
dur = (end - start)[b[:, 0] > 0]
print("per-CTA duration ns: mean %.0f min %d max %d" % (dur.mean(), dur.min(), dur.max()))
print("segments per CTA: mean %.2f min %d max %d" % (ntiles.mean(), ntiles.min(), ntiles.max()))
print("first wait (startup) ns: mean %.0f max %d" % (first_wait.mean(), first_wait.max()))
print("total wait per CTA ns: mean %.0f max %d  (%.1f%% of duration)" % (waits.mean(), waits.max(), 100 * waits.mean() / dur.mean()))
print("end-time spread ns: p10 %d p50 %d p90 %d max %d" % tuple(np.percentile(end[end > 0], [10, 50, 90, 100]).astype(int)))

# per-CTA detail: tile durations = time from end of wait i to begin of wait i+1
order = np.argsort(end)
def detail(i):
    ev = [(b[i, s], b[i, s + 1]) for s in range(2, 62, 2) if b[i, s] and b[i, s + 1]]
    durs = [ev[k + 1][0] - ev[k][1] for k in range(len(ev) - 1)]
    return "cta %3d sm %3d start %5d end %6d tiles %d waits %s tile_ns %s" % (i, b[i, 63] - 1, start[i], end[i], len(durs), [int(e[1] - e[0]) for e in ev], durs)
print("--- fastest CTAs"); [print(detail(i)) for i in order[:6]]
print("--- slowest CTAs"); [print(detail(i)) for i in order[-8:]]
sm = b[:, 63] - 1
per_sm_end = {}
for i in range(n): per_sm_end.setdefault(int(sm[i]), []).append(int(end[i]))
ends = np.array([max(v) for v in per_sm_end.values()]); cnt = np.array([len(v) for v in per_sm_end.values()])
print("SMs %d, CTAs/SM min %d max %d; per-SM last end: p10 %d p50 %d p90 %d max %d" % (len(ends), cnt.min(), cnt.max(), *np.percentile(ends, [10, 50, 90, 100]).astype(int)))
# expand-kernel phase stamps of CTA (0,0) and of the first background CTA (nctas == 1 selects that row)
eb = np.zeros((1, 64), dtype=np.uint64)
ctx.eval(P)
rc = pkg.lib().tamcmc_gpu_debug_trace(ctx.h, eb.ctypes.data_as(C.POINTER(C.c_ulonglong)), 1)
e = eb[0].astype(np.int64)
names = ["start", "params staged", "phase1 done", "passA done", "passB done", "passC done", "queue done"]
print("expand mode-CTA phases (ns since start):", [(names[i], int(e[i] - e[0])) for i in range(1, 7) if e[i]])
sub = [("ratios (warp 0)", 8), ("eta0 + scalars (warp 1)", 9), ("noise record (warp 2)", 10), ("pass A before windows (thread 5)", 11)]
print("expand sub-phases (ns since start):", [(nm, int(e[k] - e[0])) for nm, k in sub if e[k]])
print("expand background CTA: %d ns (starts %d ns after the mode CTA)" % (e[33] - e[32], e[32] - e[0]))

# per-phase cycle accounting of consumer warp 0 (trace slots 48..57), summed over the CTA's tiles
names = ["wait full", "load x,y", "fast loop", "gen loop", "background", "whittle terms", "shuffle tree", "publish partial", "combine+release", "loop top"]
ph = b[:, 48:58].astype(np.float64)
if ph.sum() > 0:
    live = ph.sum(1) > 0
    tot = ph[live].sum(1).mean()
    print("consumer warp 0, cycles per CTA (mean over %d CTAs, total %.0f = %.1f us at 1.93 GHz):" % (live.sum(), tot, tot / 1930.0))
    for k, nm in enumerate(names):
        print("  %-16s %9.0f  %5.1f %%" % (nm, ph[live, k].mean(), 100 * ph[live, k].mean() / tot))
