"""Raw dump of the -DTAMCMC_TRACE buffers of one C2 evaluation (per-CTA stamps + per-tile list sizes) for offline analysis:
gpurun_out/trace_raw.npz.  usage: TAMCMC_GPU_LIB=.../libtamcmc_gpu_trace.so python profiles/trace_dump.py [reps]"""
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
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = 1024 + 148
out = []
buf = np.zeros((n, 64), dtype=np.uint64)
ctx.eval(P)
pkg.lib().tamcmc_gpu_debug_trace(ctx.h, buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), n)
for r in range(reps):
    ctx.eval(P)
    rc = pkg.lib().tamcmc_gpu_debug_trace(ctx.h, buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), n)
    assert rc == 0
    out.append(buf.copy())
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "trace_raw.npz"), trace=np.stack(out))
print("saved", np.stack(out).shape)
