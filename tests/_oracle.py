"""ctypes wrapper around the CPU oracle (oracle/_ref/libtamcmc_oracle.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never by the product.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SO = os.path.join(_ROOT, "oracle", "_ref", "libtamcmc_oracle.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force=False):
    src = os.path.join(_ROOT, "oracle", "tamcmc_oracle.c")
    if force or not os.path.exists(_SO) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)
    ):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return _SO


def _as_d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


class Oracle:
    def __init__(self):
        build()
        L = C.CDLL(_SO)
        self.L = L
        L.orc_Pslm.restype = C.c_longdouble
        L.orc_Pslm.argtypes = [C.c_int] * 3
        L.orc_Hslm_Ritzoller1991.restype = C.c_longdouble
        L.orc_Hslm_Ritzoller1991.argtypes = [C.c_int] * 3
        L.orc_Qlm.restype = C.c_double
        L.orc_Qlm.argtypes = [C.c_int] * 2
        L.orc_amplitude_ratio.restype = None
        L.orc_amplitude_ratio.argtypes = [C.c_int, C.c_double, _dp]
        L.orc_lin_interpol.restype = C.c_double
        L.orc_lin_interpol.argtypes = [_dp, _dp, C.c_long, C.c_double]
        L.orc_linfit.restype = None
        L.orc_linfit.argtypes = [_dp, _dp, C.c_long, _dp]
        L.orc_eta0_fct.restype = C.c_double
        L.orc_eta0_fct.argtypes = [_dp, C.c_long]
        L.orc_eta0_fct_dnu.restype = C.c_double
        L.orc_eta0_fct_dnu.argtypes = [C.c_double]
        L.orc_eval_acoefs.restype = None
        L.orc_eval_acoefs.argtypes = [C.c_int, _dp, _dp]
        L.orc_set_imin_imax.restype = C.c_int
        L.orc_set_imin_imax.argtypes = [_dp, C.c_long, C.c_int] + [C.c_double] * 5 + [_ip]
        L.orc_trace_begin.restype = None
        L.orc_trace_begin.argtypes = [_ip, _ip, _ip, C.c_int]
        L.orc_trace_end.restype = C.c_int
        L.orc_build_l_mode_a1etaa3.restype = None
        L.orc_build_l_mode_a1etaa3.argtypes = [_dp, C.c_long] + [C.c_double] * 7 + [C.c_int, _dp, _dp]
        L.orc_build_l_mode_aj.restype = None
        L.orc_build_l_mode_aj.argtypes = [_dp, C.c_long] + [C.c_double] * 11 + [C.c_int, _dp, _dp]
        L.orc_likelihood_chi22p.restype = C.c_longdouble
        L.orc_likelihood_chi22p.argtypes = [_dp, _dp, C.c_long, C.c_long]
        L.orc_likelihood_chi_square.restype = C.c_longdouble
        L.orc_likelihood_chi_square.argtypes = [_dp, _dp, _dp, C.c_long]
        L.orc_call_likelihood_chi22p.restype = C.c_longdouble
        L.orc_call_likelihood_chi22p.argtypes = [_dp, _dp, C.c_long, C.c_double, C.c_double]
        self._alm_t = C.CFUNCTYPE(C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p)
        L.orc_call_model.restype = C.c_int
        L.orc_call_model.argtypes = [C.c_int, _dp, _ip, _dp, C.c_long, _dp, self._alm_t, C.c_void_p]
        L.orc_eval_chains.restype = C.c_int
        L.orc_eval_chains.argtypes = [C.c_int, _dp, C.c_int, _ip, _dp, _dp, C.c_long, C.c_int, _dp, C.c_double, _dp, C.c_int]
        L.orc_eval_chains_chi_square.restype = C.c_int
        L.orc_eval_chains_chi_square.argtypes = [C.c_int, _dp, C.c_int, _ip, _dp, _dp, _dp, C.c_long, C.c_int, _dp, _dp, C.c_int]
        L.orc_mode_table_model.restype = C.c_int
        L.orc_mode_table_model.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_long, _dp]
        L.orc_mode_table_eval_chains.restype = C.c_int
        L.orc_mode_table_eval_chains.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_long, C.c_int, _dp, C.c_double, _dp, C.c_int]
        if hasattr(L, "orc_eval_chains_fast"):
            L.orc_eval_chains_fast.restype = C.c_int
            L.orc_eval_chains_fast.argtypes = L.orc_eval_chains.argtypes

    # ---- scalars ----
    def Pslm(self, s, l, m):
        return float(self.L.orc_Pslm(s, l, m))

    def Qlm(self, l, m):
        return self.L.orc_Qlm(l, m)

    def amplitude_ratio(self, l, beta_deg):
        V = np.zeros(2 * l + 1)
        self.L.orc_amplitude_ratio(l, float(beta_deg), _p(V))
        return V

    def lin_interpol(self, x, y, x_int):
        x = _as_d(x)
        y = _as_d(y)
        return self.L.orc_lin_interpol(_p(x), _p(y), len(x), float(x_int))

    def eta0_fct(self, fl0):
        fl0 = _as_d(fl0)
        return self.L.orc_eta0_fct(_p(fl0), len(fl0))

    def set_imin_imax(self, x, l, fc, gamma, f_s, c, step):
        x = _as_d(x)
        iv = np.zeros(2, dtype=np.int32)
        rc = self.L.orc_set_imin_imax(_p(x), len(x), l, fc, gamma, f_s, c, step, iv.ctypes.data_as(_ip))
        return rc, int(iv[0]), int(iv[1])

    # ---- model / likelihood ----
    def call_model(self, model_id, params, plength, x, alm=None, trace=False):
        params = _as_d(params)
        x = _as_d(x)
        pl = np.ascontiguousarray(plength, dtype=np.int32)
        out = np.zeros(len(x))
        cb = self._alm_t(alm) if alm is not None else C.cast(None, self._alm_t)
        if trace:
            cap = 4096
            tl = np.zeros(cap, dtype=np.int32)
            t0 = np.zeros(cap, dtype=np.int32)
            t1 = np.zeros(cap, dtype=np.int32)
            self.L.orc_trace_begin(tl.ctypes.data_as(_ip), t0.ctypes.data_as(_ip), t1.ctypes.data_as(_ip), cap)
        rc = self.L.orc_call_model(model_id, _p(params), pl.ctypes.data_as(_ip), _p(x), len(x), _p(out), cb, None)
        if trace:
            n = self.L.orc_trace_end()
            return rc, out, (tl[:n].copy(), t0[:n].copy(), t1[:n].copy())
        return rc, out

    def eval_chains_chi_square(self, model_id, params, plength, x, y, sigma, Tcoefs, nthreads=0):
        params = _as_d(params)
        Nchains, Nparams = params.shape
        x, y, sigma, T = _as_d(x), _as_d(y), _as_d(sigma), _as_d(Tcoefs)
        pl = np.ascontiguousarray(plength, dtype=np.int32)
        out = np.zeros(Nchains)
        rc = self.L.orc_eval_chains_chi_square(model_id, _p(params), Nparams, pl.ctypes.data_as(_ip), _p(x), _p(y), _p(sigma), len(x),
                                               Nchains, _p(T), _p(out), int(nthreads))
        return rc, out

    def mode_table_model(self, row, Nnoise, step_mode, x, trace=False):
        row, x = _as_d(row), _as_d(x)
        out = np.zeros(len(x))
        if trace:
            cap = 8192
            tl = np.zeros(cap, dtype=np.int32)
            t0 = np.zeros(cap, dtype=np.int32)
            t1 = np.zeros(cap, dtype=np.int32)
            self.L.orc_trace_begin(tl.ctypes.data_as(_ip), t0.ctypes.data_as(_ip), t1.ctypes.data_as(_ip), cap)
        rc = self.L.orc_mode_table_model(_p(row), int(Nnoise), int(step_mode), _p(x), len(x), _p(out))
        if trace:
            n = self.L.orc_trace_end()
            return rc, out, (tl[:n].copy(), t0[:n].copy(), t1[:n].copy())
        return rc, out

    def mode_table_eval_chains(self, rows, Nnoise, step_mode, x, y, Tcoefs, p=1.0, nthreads=0):
        rows = _as_d(rows)
        Nchains, stride = rows.shape
        x, y, T = _as_d(x), _as_d(y), _as_d(Tcoefs)
        out = np.zeros(Nchains)
        rc = self.L.orc_mode_table_eval_chains(_p(rows), stride, int(Nnoise), int(step_mode), _p(x), _p(y), len(x), Nchains, _p(T),
                                               float(p), _p(out), int(nthreads))
        return rc, out

    def chi22p(self, y, model, p=1):
        y = _as_d(y)
        model = _as_d(model)
        return float(self.L.orc_likelihood_chi22p(_p(y), _p(model), len(y), int(p)))

    def call_likelihood(self, y, model, p, Tcoef):
        y = _as_d(y)
        model = _as_d(model)
        return float(self.L.orc_call_likelihood_chi22p(_p(y), _p(model), len(y), float(p), float(Tcoef)))

    def eval_chains(self, model_id, params, plength, x, y, Tcoefs, p=1.0, nthreads=0, fast=False):
        params = _as_d(params)
        Nchains, Nparams = params.shape
        x = _as_d(x)
        y = _as_d(y)
        T = _as_d(Tcoefs)
        pl = np.ascontiguousarray(plength, dtype=np.int32)
        out = np.zeros(Nchains)
        fn = self.L.orc_eval_chains_fast if fast else self.L.orc_eval_chains
        rc = fn(model_id, _p(params), Nparams, pl.ctypes.data_as(_ip), _p(x), _p(y), len(x), Nchains, _p(T), float(p), _p(out), int(nthreads))
        return rc, out


_inst = None


def get():
    global _inst
    if _inst is None:
        _inst = Oracle()
    return _inst
