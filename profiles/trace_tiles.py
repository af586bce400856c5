"""Per-tile cost fit (needs the -DTAMCMC_TRACE build): duration of every tile of consumer warp 0 against its list sizes."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g
pkg = g.load_package(); synth = pkg.synth
import bench
rng = np.random.default_rng(12345)
params, pl = synth.classic_params(rng)
x = synth.freq_axis(bench.NBINS, 500.0)
with pkg.Context(pkg.Star(3, pl, len(params), x, np.ones_like(x)), 1, [1.0]) as c0:
    M = c0.model(params)
y = synth.chi2_2dof_spectrum(rng, M)
T = synth.tcoefs(10, 1.7)
ctx = pkg.Context(pkg.Star(3, pl, len(params), x, y), 10, T)
P = ctx.pack_params([synth.perturb_chains(rng, params, pl, 10)])
for _ in range(5): ctx.eval(P)
n = 1024 + 148
buf = np.zeros((n, 64), dtype=np.uint64)
ctx.eval(P)   # (the read below also clears the buffer)
rc = pkg.lib().tamcmc_gpu_debug_trace(ctx.h, buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), n)
ctx.eval(P)
rc = pkg.lib().tamcmc_gpu_debug_trace(ctx.h, buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), n)
assert rc == 0
b = buf.astype(np.int64)
rows = []
for i in range(148):
    ev = [(b[i, s], b[i, s + 1]) for s in range(2, 48, 2) if b[i, s] and b[i, s + 1]]
    for k in range(len(ev) - 1):
        rec = int(buf[1024 + i, k])
        dur = ev[k + 1][0] - ev[k][1]
        if rec == 0 or dur <= 0 or dur > 50000: continue
        nfast, ngen, fl = rec & 0xffff, (rec >> 16) & 0xffff, (rec >> 48) & 0xff
        if not (fl & 2): continue          # one-segment tiles only
        rows.append((k, nfast, ngen, (fl >> 7) & 1, dur, ev[k][1] - ev[k][0]))
R = np.array(rows, dtype=float)
print("tiles:", len(R))
A = np.c_[np.ones(len(R)), R[:, 1], R[:, 2], R[:, 0] == 0]
coef, res, *_ = np.linalg.lstsq(A, R[:, 4], rcond=None)
print("duration_ns ~ %.0f + %.1f nfast + %.1f ngen + %.0f [first tile]" % tuple(coef))
pred = A @ coef
print("residual rms %.0f ns; duration mean %.0f" % (np.sqrt(np.mean((pred - R[:, 4]) ** 2)), R[:, 4].mean()))
for lo, hi in [(0, 0), (1, 3), (4, 8), (9, 16), (17, 64)]:
    m = (R[:, 2] >= lo) & (R[:, 2] <= hi)
    if m.sum(): print("ngen %2d..%2d: n %4d  dur mean %.0f  nfast mean %.1f  wait mean %.0f" % (lo, hi, m.sum(), R[m, 4].mean(), R[m, 1].mean(), R[m, 5].mean()))
np.save(os.path.join(ROOT, "gpurun_out", "trace_tiles.npy"), R)
