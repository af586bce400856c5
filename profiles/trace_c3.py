"""C3 (10^6 bins, 33 modes, 10 chains) with the -DTAMCMC_TRACE build: per-CTA span, background phase, ring tiles."""
import sys, os, ctypes as C, importlib.util, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as g
import _oracle
pkg = g.load_package(); synth = pkg.synth; O = _oracle.get()
spec = importlib.util.spec_from_file_location("m", os.path.join(ROOT, "tests", "golden", "make_golden_c3_c5.py"))
mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
d = tempfile.mkdtemp(); pkg.AlmGrids.make(d, 0); pkg.AlmGrids.make(d, 2); G = pkg.AlmGrids(d)
params, pl, x, y, P, T = mod.c3_inputs(synth, O, lambda l, m, t0, de, fc, user: G(l, m, t0, de, fc))
cap = int(pl[2:6].sum()); nn = int(pl[8])
rows = np.stack([pkg.expand_ajAlm(P[c], pl, cap, alm=G)[0] for c in range(10)])
ctx = pkg.Context(pkg.Star(synth.MODEL_MODE_TABLE, synth.mode_table_plength(cap, nn, 0), rows.shape[1], x, y), 10, T)
Pk = ctx.pack_params([rows])
for _ in range(5): ctx.eval(Pk)
n = 148
buf = np.zeros((n, 64), dtype=np.uint64)
ctx.eval(Pk)
pkg.lib().tamcmc_gpu_debug_trace(ctx.h, buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), n)
ctx.eval(Pk)
assert pkg.lib().tamcmc_gpu_debug_trace(ctx.h, buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), n) == 0
b = buf.astype(np.int64)
t0 = b[:, 0][b[:, 0] > 0].min()
end = b[:, 1] - t0
print("kernel span: CTA end p10 %d p50 %d max %d ns" % tuple(np.percentile(end[end > 0], [10, 50, 100]).astype(int)))
for w in (0, 1):
    s, e = b[:, 58 + 2 * w], b[:, 59 + 2 * w]
    ok = (s > 0) & (e > 0)
    if ok.sum(): print("bg phase warp %d: start %d, duration mean %d min %d max %d ns (%d CTAs)" % (w, (s[ok] - t0).mean(), (e[ok] - s[ok]).mean(), (e[ok] - s[ok]).min(), (e[ok] - s[ok]).max(), ok.sum()))
nt = []
fw = []
for i in range(n):
    ev = [(b[i, s], b[i, s + 1]) for s in range(2, 48, 2) if b[i, s] and b[i, s + 1]]
    nt.append(len(ev)); fw.append(ev[0][1] - ev[0][0] if ev else 0)
print("ring segments per CTA mean %.1f; first wait mean %d ns" % (np.mean(nt), np.mean(fw)))
ph = b[:, 48:58].astype(float)
names = ["wait full", "load x,y", "fast loop", "gen loop", "background", "whittle terms", "shuffle tree", "publish partial", "combine+release", "loop top"]
live = ph.sum(1) > 0
print("consumer warp 0 cycles:", {nm: int(ph[live, k].mean()) for k, nm in enumerate(names)})
