"""Which consumer warps of one CTA overlap which phases (needs a -DTAMCMC_TRACE -DTAMCMC_TRACE_GANTT build, see trace_gantt.sh).
For CTA 0: per sub-partition (warps w, w+4, w+8) the share of time 0/1/2/3 warps spend in the FP64-heavy phases (fast +
general loop), and a text timeline of a few tiles."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g
pkg = g.load_package(); synth = pkg.synth
import bench
rng = np.random.default_rng(12345)
params, pl = synth.classic_params(rng, trunc_c=30.0)
x = synth.freq_axis(bench.NBINS, 500.0)
with pkg.Context(pkg.Star(3, pl, len(params), x, np.ones_like(x)), 1, [1.0]) as c0:
    M = c0.model(params)
y = synth.chi2_2dof_spectrum(rng, M)
T = synth.tcoefs(10, 1.7)
ctx = pkg.Context(pkg.Star(3, pl, len(params), x, y), 10, T)
P = ctx.pack_params([synth.perturb_chains(rng, params, pl, 10)])
for _ in range(5): ctx.eval(P)
buf = np.zeros((2048, 64), dtype=np.uint64)
pkg.lib().tamcmc_gpu_debug_trace(ctx.h, buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), 2048)     # clear
ctx.eval(P)
rc = pkg.lib().tamcmc_gpu_debug_trace(ctx.h, buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), 2048)
assert rc == 0, rc
ev = buf.reshape(-1)[65536:65536 + 12 * 512].reshape(12, 512)
names = ["wait", "load", "FAST", "GEN", "bg", "whittle", "tree", "publish", "combine", "top"]
warps = []
for w in range(12):
    e = ev[w][ev[w] > 0]
    t = (e >> np.uint64(4)).astype(np.int64); k = (e & np.uint64(15)).astype(int)
    warps.append((t, k))
t0 = min(w[0][0] for w in warps if len(w[0])); t1 = max(w[0][-1] for w in warps if len(w[0]))
print("CTA 0: %d cycles = %.1f us; events per warp: %s" % (t1 - t0, (t1 - t0) / 1965.0, [len(w[0]) for w in warps]))
# phase k ENDS at its stamp: interval (previous stamp, stamp] belongs to phase k
res = 16
grid = np.full((12, (t1 - t0) // res + 2), -1, dtype=int)
for w, (t, k) in enumerate(warps):
    for i in range(1, len(t)):
        grid[w, (t[i - 1] - t0) // res:(t[i] - t0) // res] = k[i]
heavy = (grid == 2) | (grid == 3)
for sp in range(4):
    n = heavy[sp::4].sum(0)
    live = (grid[sp::4] >= 0).all(0)
    print("sub-partition %d: warps in FAST/GEN at once: " % sp + ", ".join("%d: %4.1f%%" % (j, 100 * (n[live] == j).mean()) for j in range(4))
          + "   (mean %.2f)" % n[live].mean())
for w in range(12):
    tot = (grid[w] >= 0).sum()
    print("warp %2d: " % w + " ".join("%s %4.1f%%" % (names[k], 100 * (grid[w] == k).sum() / tot) for k in range(10)))
# text timeline: one character per 128 cycles, middle of the kernel
sym = "w l F G b h t p c ."
sym = {0: "w", 1: "l", 2: "F", 3: "G", 4: "b", 5: "h", 6: "t", 7: "p", 8: "c", 9: ".", -1: " "}
a, b = int(0.35 * grid.shape[1]), int(0.35 * grid.shape[1]) + 8 * 150
for w in (0, 4, 8, 1, 5, 9):
    print("warp %2d |" % w + "".join(sym[int(v)] for v in grid[w, a:b:8]))
