"""tamcmc-c_b200: B200-native hot path of TAMCMC-C (model spectrum + Whittle logL).

This Python module is only the ctypes binding of the C ABI in include/tamcmc_gpu.h
(used by tests/, bench.py and __graft_entry__.py).  The product is the shared library
`libtamcmc_gpu.so` built from csrc/ (hand-written sm_100a CUDA); the reference-facing
host mirror is the C++ header host/model_def_gpu.hpp.  There is no CPU fallback: if the
library is missing or no CUDA device is usable, every call raises.

The directory name has a hyphen, so import it through `__graft_entry__.load_package()`
(module alias `tamcmc_c_b200`).
"""
import ctypes as C
import os

import numpy as np

from . import synth  # noqa: F401  (synthetic inputs; numpy only)
from . import formats  # noqa: F401  (readers for the reference's simplest on-disk inputs; numpy only)
from . import model_setup  # noqa: F401  (parsed .model file -> parameter vector, plength, relax flags, prior table; numpy only)

_HERE = os.path.dirname(os.path.abspath(__file__))
# TAMCMC_GPU_LIB selects another build of the SAME CUDA library (e.g. the -DTAMCMC_TRACE profiling build)
LIB_PATH = os.environ.get("TAMCMC_GPU_LIB") or os.path.join(_HERE, "libtamcmc_gpu.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_ucp = C.POINTER(C.c_ubyte)

OK, ERR_ARG, ERR_MODEL, ERR_CUDA, ERR_WINDOW, ERR_NONFINITE, ERR_LIKELIHOOD = range(7)
CHAIN_WINDOW, CHAIN_NONFINITE, CHAIN_BADCFG, CHAIN_INACTIVE = 1, 2, 4, 8

# every symbol include/tamcmc_gpu.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "tamcmc_gpu_create", "tamcmc_gpu_destroy", "tamcmc_gpu_eval", "tamcmc_gpu_eval_begin", "tamcmc_gpu_params_staging", "tamcmc_gpu_eval_end", "tamcmc_gpu_eval_device", "tamcmc_gpu_pt_swap_device",
    "tamcmc_gpu_sync", "tamcmc_gpu_model", "tamcmc_gpu_windows", "tamcmc_gpu_components",
    "tamcmc_gpu_params_stride", "tamcmc_gpu_nstars", "tamcmc_gpu_nchains", "tamcmc_gpu_pairs_last",
    "tamcmc_gpu_set_profiling", "tamcmc_gpu_get_kernel_ms", "tamcmc_gpu_launch_count",
    "tamcmc_gpu_debug_trace", "tamcmc_gpu_fp64_peak", "tamcmc_gpu_strerror", "tamcmc_gpu_last_error", "tamcmc_gpu_abi_version",
    "tamcmc_gpu_exchange_create", "tamcmc_gpu_exchange_attach", "tamcmc_gpu_exchange_attach_ptrs", "tamcmc_gpu_exchange_buffer",
    "tamcmc_host_alm", "tamcmc_host_expand_ajAlm",
    "tamcmc_host_expand_rgb_v4", "tamcmc_host_armm_solve_from_l0", "tamcmc_host_armm_solve_O2p", "tamcmc_host_spline_eval",
    "tamcmc_gpu_rgb_create", "tamcmc_gpu_rgb_destroy", "tamcmc_gpu_rgb_expand", "tamcmc_gpu_rgb_timings", "tamcmc_gpu_rgb_counts", "tamcmc_gpu_rgb_last_error",
    "tamcmc_host_rgb_expand_emulated",
    "tamcmc_alm_grids_load", "tamcmc_alm_grids_free", "tamcmc_alm_grids_eval", "tamcmc_alm_grids_shape", "tamcmc_alm_grids_nodes",
    "tamcmc_alm_grids_make", "tamcmc_alm_grids_last_error",
]

ALM_FN = C.CFUNCTYPE(C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p)


class StarStruct(C.Structure):
    _fields_ = [
        ("model_id", C.c_int),
        ("plength", C.c_int * 11),
        ("Nparams", C.c_int),
        ("x", _dp),
        ("y", _dp),
        ("N", C.c_long),
        ("N_global", C.c_long),
        ("bin_offset", C.c_long),
        ("x_first", C.c_double),
        ("x_second", C.c_double),
        ("x_last", C.c_double),
        ("sigma_y", _dp),
    ]


class TamcmcError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(msg)
        self.status = status


_lib = None


def lib():
    """Load libtamcmc_gpu.so (fails loudly: there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TamcmcError(ERR_CUDA, "libtamcmc_gpu.so not built: run __graft_entry__.build() (no CPU fallback exists)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.tamcmc_gpu_create.restype = C.c_int
    L.tamcmc_gpu_create.argtypes = [C.c_int, C.c_int, C.POINTER(StarStruct), C.c_int, _dp, C.c_double, C.c_int, C.POINTER(vp)]
    L.tamcmc_gpu_destroy.restype = None
    L.tamcmc_gpu_destroy.argtypes = [vp]
    L.tamcmc_gpu_eval.restype = C.c_int
    L.tamcmc_gpu_eval.argtypes = [vp, _dp, _ucp, _dp, _ip]
    L.tamcmc_gpu_eval_begin.restype = C.c_int
    L.tamcmc_gpu_eval_begin.argtypes = [vp, _dp, _ucp]
    L.tamcmc_gpu_eval_end.restype = C.c_int
    L.tamcmc_gpu_eval_end.argtypes = [vp, _dp, _ip]
    L.tamcmc_gpu_pt_swap_device.restype = C.c_int
    L.tamcmc_gpu_pt_swap_device.argtypes = [vp, C.c_int, C.c_int, C.c_double, vp, vp, vp, vp, vp]
    L.tamcmc_gpu_eval_device.restype = C.c_int
    L.tamcmc_gpu_eval_device.argtypes = [vp, vp, vp, vp, C.c_int, vp]
    L.tamcmc_gpu_sync.restype = C.c_int
    L.tamcmc_gpu_sync.argtypes = [vp]
    L.tamcmc_gpu_model.restype = C.c_int
    L.tamcmc_gpu_model.argtypes = [vp, C.c_int, _dp, _dp]
    L.tamcmc_gpu_windows.restype = C.c_int
    L.tamcmc_gpu_windows.argtypes = [vp, C.c_int, _dp, C.c_int, _ip, _ip, _ip, _ip]
    L.tamcmc_gpu_components.restype = C.c_int
    L.tamcmc_gpu_components.argtypes = [vp, C.c_int, _dp, C.c_int, _ip, _ip, _ip, _dp, _dp, _dp]
    for f in ("tamcmc_gpu_params_stride", "tamcmc_gpu_nstars", "tamcmc_gpu_nchains"):
        getattr(L, f).restype = C.c_int
        getattr(L, f).argtypes = [vp]
    L.tamcmc_gpu_pairs_last.restype = C.c_long
    L.tamcmc_gpu_pairs_last.argtypes = [vp]
    L.tamcmc_gpu_set_profiling.restype = C.c_int
    L.tamcmc_gpu_set_profiling.argtypes = [vp, C.c_int]
    L.tamcmc_gpu_get_kernel_ms.restype = C.c_int
    L.tamcmc_gpu_get_kernel_ms.argtypes = [vp, C.POINTER(C.c_long), _dp, _dp]
    L.tamcmc_gpu_launch_count.restype = C.c_long
    L.tamcmc_gpu_launch_count.argtypes = [vp]
    L.tamcmc_gpu_debug_trace.restype = C.c_int
    L.tamcmc_gpu_debug_trace.argtypes = [vp, C.POINTER(C.c_ulonglong), C.c_int]
    L.tamcmc_gpu_fp64_peak.restype = C.c_int
    L.tamcmc_gpu_fp64_peak.argtypes = [C.c_int, _dp]
    L.tamcmc_gpu_strerror.restype = C.c_char_p
    L.tamcmc_gpu_strerror.argtypes = [C.c_int]
    L.tamcmc_gpu_last_error.restype = C.c_char_p
    L.tamcmc_gpu_abi_version.restype = C.c_int
    L.tamcmc_gpu_exchange_create.restype = C.c_int
    L.tamcmc_gpu_exchange_create.argtypes = [vp, vp]
    L.tamcmc_gpu_exchange_attach.restype = C.c_int
    L.tamcmc_gpu_exchange_attach.argtypes = [vp, C.c_int, C.c_int, vp]
    L.tamcmc_gpu_exchange_attach_ptrs.restype = C.c_int
    L.tamcmc_gpu_exchange_attach_ptrs.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.tamcmc_gpu_exchange_buffer.restype = vp
    L.tamcmc_gpu_exchange_buffer.argtypes = [vp]
    L.tamcmc_host_alm.restype = C.c_double
    L.tamcmc_host_alm.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
    L.tamcmc_host_expand_ajAlm.restype = C.c_int
    L.tamcmc_host_expand_ajAlm.argtypes = [_dp, _ip, ALM_FN, vp, C.c_int, _dp, _ip]
    L.tamcmc_host_expand_rgb_v4.restype = C.c_int
    L.tamcmc_host_expand_rgb_v4.argtypes = [C.c_int, _dp, _ip, C.c_double, C.c_int, _dp, _ip]
    L.tamcmc_gpu_rgb_create.restype = C.c_int
    L.tamcmc_gpu_rgb_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int]
    L.tamcmc_gpu_rgb_destroy.restype = None
    L.tamcmc_gpu_rgb_destroy.argtypes = [C.c_void_p]
    L.tamcmc_gpu_rgb_expand.restype = C.c_int
    L.tamcmc_gpu_rgb_expand.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, _ip, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_int, _ip, _ip, _ip]
    L.tamcmc_gpu_rgb_timings.restype = None
    L.tamcmc_gpu_rgb_timings.argtypes = [C.c_void_p, _dp]
    L.tamcmc_gpu_rgb_last_error.restype = C.c_char_p
    L.tamcmc_gpu_rgb_counts.restype = None
    L.tamcmc_gpu_rgb_counts.argtypes = [C.c_void_p, C.POINTER(C.c_long), C.POINTER(C.c_long)]
    L.tamcmc_host_rgb_expand_emulated.restype = C.c_int
    L.tamcmc_host_rgb_expand_emulated.argtypes = [C.c_int, _dp, _ip, C.c_double, C.c_int, _dp, _ip, C.c_int, _ip]
    L.tamcmc_host_armm_solve_from_l0.restype = C.c_int
    L.tamcmc_host_armm_solve_from_l0.argtypes = [_dp, C.c_int, C.c_int] + [C.c_double] * 7 + [C.c_int, _dp, _ip, _dp, _dp, _ip, _dp, _ip]
    L.tamcmc_gpu_params_staging.restype = _dp
    L.tamcmc_gpu_params_staging.argtypes = [vp, _ip]
    L.tamcmc_host_armm_solve_O2p.restype = C.c_int
    L.tamcmc_host_armm_solve_O2p.argtypes = [C.c_double, C.c_double, C.c_int] + [C.c_double] * 9 + [C.c_int, _dp, _ip, _dp, _dp, _ip, _dp, _ip]
    L.tamcmc_host_spline_eval.restype = C.c_int
    L.tamcmc_host_spline_eval.argtypes = [_dp, _dp, C.c_int, C.c_int, _dp, C.c_int, _dp]
    L.tamcmc_alm_grids_load.restype = C.c_int
    L.tamcmc_alm_grids_load.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.tamcmc_alm_grids_free.restype = None
    L.tamcmc_alm_grids_free.argtypes = [vp]
    L.tamcmc_alm_grids_eval.restype = C.c_double
    L.tamcmc_alm_grids_eval.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, vp]
    L.tamcmc_alm_grids_shape.restype = C.c_int
    L.tamcmc_alm_grids_shape.argtypes = [vp, C.c_int, C.c_int, C.c_int, _ip, _ip]
    L.tamcmc_alm_grids_nodes.restype = C.c_int
    L.tamcmc_alm_grids_nodes.argtypes = [vp, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp]
    L.tamcmc_alm_grids_make.restype = C.c_int
    L.tamcmc_alm_grids_make.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]
    L.tamcmc_alm_grids_last_error.restype = C.c_char_p
    _lib = L
    return L


def _raise(rc):
    L = lib()
    msg = L.tamcmc_gpu_strerror(rc).decode()
    if rc == ERR_CUDA:
        msg += " -- " + L.tamcmc_gpu_last_error().decode()
    raise TamcmcError(rc, msg)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Star:
    """Host-side description of one star/slice (reference: Data{x,y,Nx} + model selection)."""

    def __init__(self, model_id, plength, Nparams, x, y, N_global=0, bin_offset=0, x_first=0.0, x_second=0.0, x_last=0.0, sigma_y=None):
        self.model_id = int(model_id)
        self.plength = np.ascontiguousarray(plength, dtype=np.int32)
        self.Nparams = int(Nparams)
        self.x = _d(x)
        self.y = _d(y)
        self.N_global, self.bin_offset = int(N_global), int(bin_offset)
        self.x_first, self.x_second, self.x_last = float(x_first), float(x_second), float(x_last)
        self.sigma_y = None if sigma_y is None else _d(sigma_y)     # chi_square likelihood only

    @classmethod
    def shard(cls, model_id, plength, Nparams, x, y, lo, hi, sigma_y=None):
        """Bins [lo,hi) of a spectrum (bin-sharding over GPUs)."""
        x = _d(x)
        return cls(model_id, plength, Nparams, x[lo:hi], _d(y)[lo:hi], N_global=len(x), bin_offset=lo,
                   x_first=x[0], x_second=x[1], x_last=x[-1], sigma_y=None if sigma_y is None else _d(sigma_y)[lo:hi])

    def struct(self):
        s = StarStruct()
        s.model_id = self.model_id
        for i in range(11):
            s.plength[i] = int(self.plength[i])
        s.Nparams = self.Nparams
        s.x = self.x.ctypes.data_as(_dp)
        s.y = self.y.ctypes.data_as(_dp)
        s.N = len(self.x)
        s.N_global, s.bin_offset = self.N_global, self.bin_offset
        s.x_first, s.x_second, s.x_last = self.x_first, self.x_second, self.x_last
        s.sigma_y = self.sigma_y.ctypes.data_as(_dp) if self.sigma_y is not None else None
        return s


class Context:
    """Thin owner of a tamcmc_gpu_ctx.  Mirrors the hot-path half of the reference's Model_def:
    eval() = generate_model's call_model + call_likelihood for all chains (model_def.cpp:466-482),
    model() = call_model_explicit (model_def.cpp:209-218)."""

    def __init__(self, stars, Nchains, Tcoefs, p=1.0, likelihood_id=0, device=0):
        if isinstance(stars, Star):
            stars = [stars]
        self.stars = list(stars)
        self.Nchains = int(Nchains)
        L = lib()
        arr = (StarStruct * len(self.stars))(*[s.struct() for s in self.stars])
        T = _d(Tcoefs)
        if len(T) != self.Nchains:
            raise TamcmcError(ERR_ARG, "len(Tcoefs) != Nchains")
        h = C.c_void_p()
        rc = L.tamcmc_gpu_create(int(device), len(self.stars), arr, self.Nchains, T.ctypes.data_as(_dp), float(p), int(likelihood_id), C.byref(h))
        if rc != OK:
            _raise(rc)
        self.h = h
        self.nstars = len(self.stars)
        self.params_stride = L.tamcmc_gpu_params_stride(h)

    def close(self):
        if getattr(self, "h", None):
            lib().tamcmc_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def pack_params(self, params):
        """params: array [nstars, Nchains, Nparams_s] (or [Nchains, Nparams] for one star, or a list
        of per-star arrays) -> contiguous [nstars, Nchains, params_stride]."""
        if isinstance(params, np.ndarray) and params.ndim == 2:
            params = [params]
        out = np.zeros((self.nstars, self.Nchains, self.params_stride))
        for s in range(self.nstars):
            p = _d(params[s])
            out[s, :, : p.shape[1]] = p
        return out

    def eval(self, params, active=None, raise_on_error=True):
        """One batched evaluation from HOST buffers: returns (logL[nstars, Nchains], status)."""
        P = params if (isinstance(params, np.ndarray) and params.ndim == 3 and params.shape[2] == self.params_stride
                       and params.flags.c_contiguous and params.dtype == np.float64) else self.pack_params(params)
        out = np.empty((self.nstars, self.Nchains))
        st = np.empty((self.nstars, self.Nchains), dtype=np.int32)
        am = None
        if active is not None:
            am_arr = np.ascontiguousarray(active, dtype=np.uint8).reshape(self.nstars, self.Nchains)
            am = am_arr.ctypes.data_as(_ucp)
        rc = lib().tamcmc_gpu_eval(self.h, P.ctypes.data_as(_dp), am, out.ctypes.data_as(_dp), st.ctypes.data_as(_ip))
        if rc != OK and raise_on_error:
            _raise(rc)
        self.last_rc = rc
        return out, st

    def params_staging(self):
        """The context's pinned staging block as a numpy view [nstars, Nchains, params_stride]: rows written here and passed to
        eval() / bind_host_buffers() skip the host-side copy of the call (tamcmc_gpu_params_staging)."""
        st = C.c_int(0)
        ptr = lib().tamcmc_gpu_params_staging(self.h, C.byref(st))
        if not ptr:
            _raise(ERR_ARG)
        n = self.nstars * self.Nchains * st.value
        return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(self.nstars, self.Nchains, st.value)

    def bind_host_buffers(self, params, logL_out, status_out=None, active=None):
        """Pre-convert caller-owned contiguous host arrays to C pointers for repeated evaluations: returns a zero-argument
        callable that runs tamcmc_gpu_eval on them and returns its status code (what a C/C++ caller's loop does; skips the
        per-call numpy/ctypes conversions of eval())."""
        assert params.dtype == np.float64 and params.flags.c_contiguous and params.size == self.nstars * self.Nchains * self.params_stride
        assert logL_out.dtype == np.float64 and logL_out.flags.c_contiguous and logL_out.size == self.nstars * self.Nchains
        fn = lib().tamcmc_gpu_eval
        h = self.h
        pp = params.ctypes.data_as(_dp)
        po = logL_out.ctypes.data_as(_dp)
        ps = status_out.ctypes.data_as(_ip) if status_out is not None else None
        pa = active.ctypes.data_as(_ucp) if active is not None else None
        keep = (params, logL_out, status_out, active)

        def call(_keep=keep):
            return fn(h, pp, pa, po, ps)
        return call

    def pt_swap_device(self, A, u, d_params_ptr, d_logL_ptr, d_logPrior_ptr=None, d_swapped_ptr=None, star=0, stream=None):
        """Parallel-tempering swap of chains A, A+1 decided and applied on the device (MALA.cpp:397-461)."""
        rc = lib().tamcmc_gpu_pt_swap_device(self.h, int(star), int(A), float(u), C.c_void_p(d_params_ptr), C.c_void_p(d_logL_ptr),
                                             C.c_void_p(d_logPrior_ptr) if d_logPrior_ptr else None,
                                             C.c_void_p(d_swapped_ptr) if d_swapped_ptr else None, C.c_void_p(stream) if stream else None)
        if rc != OK:
            _raise(rc)

    def eval_device(self, d_params_ptr, d_logL_ptr, d_active_ptr=None, raw_sum=False, stream=None):
        rc = lib().tamcmc_gpu_eval_device(self.h, C.c_void_p(d_params_ptr), C.c_void_p(d_active_ptr) if d_active_ptr else None,
                                          C.c_void_p(d_logL_ptr), 1 if raw_sum else 0, C.c_void_p(stream) if stream else None)
        if rc != OK:
            _raise(rc)

    def exchange_handle(self):
        """Allocates this rank's exchange buffer of a bin-sharded spectrum; returns its 64-byte CUDA IPC handle (numpy uint8)."""
        h = np.zeros(64, dtype=np.uint8)
        rc = lib().tamcmc_gpu_exchange_create(self.h, h.ctypes.data_as(C.c_void_p))
        if rc != OK:
            _raise(rc)
        return h

    def exchange_attach(self, rank, world, handles):
        """handles: uint8 [world, 64] in rank order (all_gather of exchange_handle()).  Afterwards eval / eval_device
        return the log-likelihood of the WHOLE spectrum on every rank (peer-memory exchange inside the fused kernel)."""
        hb = np.ascontiguousarray(handles, dtype=np.uint8).reshape(world, 64)
        rc = lib().tamcmc_gpu_exchange_attach(self.h, int(rank), int(world), hb.ctypes.data_as(C.c_void_p))
        if rc != OK:
            _raise(rc)

    def sync(self):
        rc = lib().tamcmc_gpu_sync(self.h)
        if rc != OK:
            _raise(rc)

    def model(self, params_row, star=0):
        row = _d(params_row)
        out = np.empty(len(self.stars[star].x))
        rc = lib().tamcmc_gpu_model(self.h, star, row.ctypes.data_as(_dp), out.ctypes.data_as(_dp))
        if rc != OK:
            _raise(rc)
        return out

    def windows(self, params_row, star=0, cap=8192, raise_on_error=True):
        row = _d(params_row)
        n = C.c_int(0)
        l = np.zeros(cap, dtype=np.int32)
        i0 = np.zeros(cap, dtype=np.int32)
        i1 = np.zeros(cap, dtype=np.int32)
        rc = lib().tamcmc_gpu_windows(self.h, star, row.ctypes.data_as(_dp), cap, C.byref(n), l.ctypes.data_as(_ip),
                                      i0.ctypes.data_as(_ip), i1.ctypes.data_as(_ip))
        if rc != OK and raise_on_error:
            _raise(rc)
        k = min(n.value, cap)
        return rc, l[:k].copy(), i0[:k].copy(), i1[:k].copy()

    def components(self, params_row, star=0, cap=65536):
        row = _d(params_row)
        n = C.c_int(0)
        mi = np.zeros(cap, dtype=np.int32)
        m = np.zeros(cap, dtype=np.int32)
        nu = np.zeros(cap)
        h = np.zeros(cap)
        w = np.zeros(cap)
        rc = lib().tamcmc_gpu_components(self.h, star, row.ctypes.data_as(_dp), cap, C.byref(n), mi.ctypes.data_as(_ip),
                                         m.ctypes.data_as(_ip), nu.ctypes.data_as(_dp), h.ctypes.data_as(_dp), w.ctypes.data_as(_dp))
        if rc != OK:
            _raise(rc)
        k = min(n.value, cap)
        return mi[:k].copy(), m[:k].copy(), nu[:k].copy(), h[:k].copy(), w[:k].copy()

    def pairs_last(self):
        return int(lib().tamcmc_gpu_pairs_last(self.h))

    def set_profiling(self, on=True):
        lib().tamcmc_gpu_set_profiling(self.h, 1 if on else 0)

    def kernel_ms(self):
        n = C.c_long(0)
        a = C.c_double(0)
        b = C.c_double(0)
        lib().tamcmc_gpu_get_kernel_ms(self.h, C.byref(n), C.byref(a), C.byref(b))
        return n.value, a.value, b.value

    def launch_count(self):
        return int(lib().tamcmc_gpu_launch_count(self.h))


def host_alm(l, m, theta0, delta, filter_code=0):
    """Alm(l, m, theta0, delta) [radians] of the gate (0) / triangle (2) filter (activity.cpp:221-246)."""
    return lib().tamcmc_host_alm(int(l), int(m), float(theta0), float(delta), int(filter_code))


class AlmGrids:
    """The reference's precomputed Alm grids + GSL-style bicubic interpolation (Config::Config, config.cpp:77-147;
    Alm_interp_iter_preinitialised, Alm_interpol.cpp:188-348)."""

    def __init__(self, grid_dir):
        h = C.c_void_p()
        rc = lib().tamcmc_alm_grids_load(os.fsencode(grid_dir), C.byref(h))
        if rc != OK:
            raise TamcmcError(rc, lib().tamcmc_alm_grids_last_error().decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().tamcmc_alm_grids_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __call__(self, l, m, theta0, delta, filter_code=0):
        return lib().tamcmc_alm_grids_eval(int(l), int(m), float(theta0), float(delta), int(filter_code), self.h)

    def nodes(self, l, m, filter_code=0):
        nx, ny = C.c_int(0), C.c_int(0)
        rc = lib().tamcmc_alm_grids_shape(self.h, int(filter_code), int(l), int(m), C.byref(nx), C.byref(ny))
        if rc != OK:
            _raise(rc)
        x, y, z = np.zeros(nx.value), np.zeros(ny.value), np.zeros((ny.value, nx.value))
        lib().tamcmc_alm_grids_nodes(self.h, int(filter_code), int(l), int(m), x.ctypes.data_as(_dp), y.ctypes.data_as(_dp), z.ctypes.data_as(_dp))
        return x, y, z

    @staticmethod
    def make(out_dir, filter_code, lmax=3, resol=np.pi / 180., theta=(0.0, np.pi / 2), delta=(0.0, np.pi / 4)):
        """GridMaker (do_grids.cpp, make_grids.cpp): defaults = the ranges of the reference's shipped 1-degree grids."""
        rc = lib().tamcmc_alm_grids_make(os.fsencode(out_dir), int(filter_code), int(lmax), float(resol), float(theta[0]), float(theta[1]),
                                         float(delta[0]), float(delta[1]))
        if rc != OK:
            raise TamcmcError(rc, lib().tamcmc_alm_grids_last_error().decode())


def expand_ajAlm(params, plength, capacity, alm=None):
    """Host expander of model_MS_Global_ajAlm_HarveyLike (models.cpp:1411-1746): -> (mode-table row, nmodes).
    alm: None (direct integral tamcmc_host_alm), an AlmGrids (the reference's grid interpolation), or a Python callable."""
    p = _d(params)
    pl = np.ascontiguousarray(plength, dtype=np.int32)
    row = np.zeros(synth.mode_table_nparams(capacity, int(pl[8])))
    n = C.c_int(0)
    user = None
    if isinstance(alm, AlmGrids):
        cb = C.cast(lib().tamcmc_alm_grids_eval, ALM_FN)
        user = alm.h
    else:
        cb = ALM_FN(alm) if alm is not None else C.cast(None, ALM_FN)
    rc = lib().tamcmc_host_expand_ajAlm(p.ctypes.data_as(_dp), pl.ctypes.data_as(_ip), cb, user, int(capacity), row.ctypes.data_as(_dp), C.byref(n))
    if rc != OK:
        _raise(rc)
    return row, n.value


def expand_rgb_v4(model_id, params, plength, step, capacity):
    """Host expander of model_RGB_asympt_aj_{AppWidth (25), CteWidth (27)}_HarveyLike_v4 (models.cpp:4684-5079, 4334-4682):
    -> (mode-table row, nmodes).  step = x[2] - x[1]."""
    p = _d(params)
    pl = np.ascontiguousarray(plength, dtype=np.int32)
    row = np.zeros(synth.mode_table_nparams(capacity, int(pl[8])))
    n = C.c_int(0)
    rc = lib().tamcmc_host_expand_rgb_v4(int(model_id), p.ctypes.data_as(_dp), pl.ctypes.data_as(_ip), float(step), int(capacity),
                                         row.ctypes.data_as(_dp), C.byref(n))
    if rc != OK:
        _raise(rc)
    return row, n.value


def expand_rgb_v4_emulated(model_id, params, plength, step, capacity, exact_trig=1):
    """TEST HOOK: the device solver's decomposition (csrc/rgb_solver.cuh) run on the host -> (row, nmodes, flags)."""
    p = _d(params)
    pl = np.ascontiguousarray(plength, dtype=np.int32)
    row = np.zeros(synth.mode_table_nparams(capacity, int(pl[8])))
    n, fl = C.c_int(0), C.c_int(0)
    rc = lib().tamcmc_host_rgb_expand_emulated(int(model_id), p.ctypes.data_as(_dp), pl.ctypes.data_as(_ip), float(step), int(capacity),
                                               row.ctypes.data_as(_dp), C.byref(n), int(exact_trig), C.byref(fl))
    if rc != OK:
        _raise(rc)
    return row, n.value, fl.value


class RgbExpander:
    """tamcmc_gpu_rgb_*: the red-giant expander with the mixed-mode pair loop and the zeta normalisation on the device, all chains of a
    step in one call.  expand(params[nchains][nparams]) -> (rows, nmodes, status, path); rows may be a caller-owned array (e.g. a view of
    Context.params_staging()) so that the evaluation reads them in place."""

    def __init__(self, model_id, plength, step, capacity, max_chains, device=0):
        self.model_id, self.step, self.capacity, self.max_chains = int(model_id), float(step), int(capacity), int(max_chains)
        self.pl = np.ascontiguousarray(plength, dtype=np.int32)
        self.row_len = synth.mode_table_nparams(self.capacity, int(self.pl[8]))
        self.h = C.c_void_p()
        rc = lib().tamcmc_gpu_rgb_create(C.byref(self.h), int(device), self.max_chains)
        if rc != OK:
            raise TamcmcError(rc, (lib().tamcmc_gpu_rgb_last_error() or b"").decode())

    def expand(self, params, rows_out=None):
        P = np.ascontiguousarray(params, dtype=np.float64)
        if P.ndim == 1:
            P = P[None, :]
        n = P.shape[0]
        rows = np.zeros((n, self.row_len)) if rows_out is None else rows_out
        assert rows.dtype == np.float64 and rows.shape[0] >= n and rows.strides[1] == 8
        nm = np.zeros(n, dtype=np.int32)
        st = np.zeros(n, dtype=np.int32)
        path = np.zeros(n, dtype=np.int32)
        rc = lib().tamcmc_gpu_rgb_expand(self.h, self.model_id, P.ctypes.data, P.strides[0] // 8, self.pl.ctypes.data_as(_ip), self.step, n,
                                         self.capacity, rows.ctypes.data, rows.strides[0] // 8, nm.ctypes.data_as(_ip), st.ctypes.data_as(_ip),
                                         path.ctypes.data_as(_ip))
        if rc != OK:
            raise TamcmcError(rc, (lib().tamcmc_gpu_rgb_last_error() or b"").decode())
        return rows, nm, st, path

    def timings(self):
        t = np.zeros(4)
        lib().tamcmc_gpu_rgb_timings(self.h, t.ctypes.data_as(_dp))
        return dict(prepare_ms=t[0], device_ms=t[1], finish_ms=t[2], total_ms=t[3])

    def counts(self):
        """(chain set-ups asked for, of which handed to the host solver) since this expander was created"""
        a, b = C.c_long(0), C.c_long(0)
        lib().tamcmc_gpu_rgb_counts(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def close(self):
        if self.h:
            lib().tamcmc_gpu_rgb_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def armm_solve_from_l0(nu_l0, el, delta0l, DPl, alpha, q, resol, freq_min, freq_max, cap=4096):
    """solve_mm_asymptotic_O2from_l0 (solver_mm.cpp:624-746): -> nu_m, nu_p, dnup, nu_g."""
    f = _d(nu_l0)
    out = [np.zeros(cap) for _ in range(4)]
    n = [C.c_int(0) for _ in range(3)]
    rc = lib().tamcmc_host_armm_solve_from_l0(f.ctypes.data_as(_dp), len(f), int(el), float(delta0l), float(DPl), float(alpha), float(q), float(resol),
                                              float(freq_min), float(freq_max), cap, out[0].ctypes.data_as(_dp), C.byref(n[0]), out[1].ctypes.data_as(_dp),
                                              out[2].ctypes.data_as(_dp), C.byref(n[1]), out[3].ctypes.data_as(_dp), C.byref(n[2]))
    if rc != OK:
        _raise(rc)
    return out[0][:n[0].value].copy(), out[1][:n[1].value].copy(), out[2][:n[1].value].copy(), out[3][:n[2].value].copy()


def armm_solve_O2p(Dnu_p, epsilon, el, delta0l, alpha_p, nmax, DPl, alpha, q, fmin, fmax, resol, cap=4096):
    """solve_mm_asymptotic_O2p (solver_mm.cpp:470-604): -> nu_m, nu_p, dnup, nu_g."""
    out = [np.zeros(cap) for _ in range(4)]
    n = [C.c_int(0) for _ in range(3)]
    rc = lib().tamcmc_host_armm_solve_O2p(float(Dnu_p), float(epsilon), int(el), float(delta0l), float(alpha_p), float(nmax), float(DPl), float(alpha),
                                          float(q), float(fmin), float(fmax), float(resol), cap, out[0].ctypes.data_as(_dp), C.byref(n[0]),
                                          out[1].ctypes.data_as(_dp), out[2].ctypes.data_as(_dp), C.byref(n[1]), out[3].ctypes.data_as(_dp), C.byref(n[2]))
    if rc != OK:
        _raise(rc)
    return out[0][:n[0].value].copy(), out[1][:n[1].value].copy(), out[2][:n[1].value].copy(), out[3][:n[2].value].copy()


def spline_eval(x, y, xq, kind=1):
    """tk::spline with zero second derivatives at the ends (spline.h): kind 1 cubic, 2 Hermite."""
    x, y, xq = _d(x), _d(y), _d(np.atleast_1d(xq))
    out = np.zeros(len(xq))
    rc = lib().tamcmc_host_spline_eval(x.ctypes.data_as(_dp), y.ctypes.data_as(_dp), len(x), int(kind), xq.ctypes.data_as(_dp), len(xq), out.ctypes.data_as(_dp))
    if rc != OK:
        _raise(rc)
    return out


def fp64_peak(device=0):
    v = C.c_double(0)
    rc = lib().tamcmc_gpu_fp64_peak(int(device), C.byref(v))
    if rc != OK:
        _raise(rc)
    return v.value
