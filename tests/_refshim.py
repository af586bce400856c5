"""ctypes wrapper around oracle/_ref/libtamcmc_refshim.so: the REFERENCE's own hot-path sources compiled against the
Eigen-API shim (oracle/Makefile `ref`).  Test infrastructure only.  Available where /root/reference was present
at build time (the build container); the built .so travels to the GPU box with the snapshot."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(os.path.dirname(_HERE), "oracle", "_ref", "libtamcmc_refshim.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


SO_O3 = os.path.join(os.path.dirname(_HERE), "oracle", "_ref", "libtamcmc_refshim_O3.so")   # bench arm (-O3 like the reference's release build)


def available():
    return os.path.exists(SO)


def available_O3():
    return os.path.exists(SO_O3)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


class Ref:
    def __init__(self, so=SO):
        L = C.CDLL(so)
        self.L = L
        L.ref_Pslm.restype = C.c_longdouble
        L.ref_Pslm.argtypes = [C.c_int] * 3
        L.ref_Qlm.restype = C.c_double
        L.ref_Qlm.argtypes = [C.c_int] * 2
        L.ref_amplitude_ratio.argtypes = [C.c_int, C.c_double, _dp]
        L.ref_lin_interpol.restype = C.c_double
        L.ref_lin_interpol.argtypes = [_dp, _dp, C.c_long, C.c_double]
        L.ref_eta0_fct.restype = C.c_double
        L.ref_eta0_fct.argtypes = [_dp, C.c_long]
        L.ref_set_imin_imax.argtypes = [_dp, C.c_long, C.c_int] + [C.c_double] * 5 + [_ip]
        L.ref_build_l_mode_a1etaa3.argtypes = [_dp, C.c_long] + [C.c_double] * 7 + [C.c_int, _dp, _dp]
        L.ref_build_l_mode_aj.argtypes = [_dp, C.c_long] + [C.c_double] * 11 + [C.c_int, _dp, _dp]
        L.ref_harvey_like.argtypes = [_dp, C.c_int, _dp, _dp, C.c_long, C.c_int, _dp]
        L.ref_likelihood_chi22p.restype = C.c_longdouble
        L.ref_likelihood_chi22p.argtypes = [_dp, _dp, C.c_long, C.c_long]
        L.ref_call_model.restype = C.c_int
        L.ref_call_model.argtypes = [C.c_int, _dp, C.c_int, _ip, _dp, C.c_long, _dp]
        L.ref_call_model_recorded.restype = C.c_int
        L.ref_call_model_recorded.argtypes = [C.c_int, _dp, C.c_int, _ip, _dp, C.c_long, _dp, _dp, C.c_int, _ip]
        L.ref_eval_chains.restype = C.c_int
        L.ref_eval_chains.argtypes = [C.c_int, _dp, C.c_int, _ip, _dp, _dp, C.c_long, C.c_int, _dp, C.c_double, _dp, C.c_int]
        L.ref_max_threads.restype = C.c_int
        L.ref_likelihood_chi_square.restype = C.c_longdouble
        L.ref_likelihood_chi_square.argtypes = [_dp, _dp, _dp, C.c_long]
        if hasattr(L, "ref_solve_mm_from_l0"):
            L.ref_solve_mm_from_l0.restype = C.c_int
            L.ref_solve_mm_from_l0.argtypes = [_dp, C.c_int, C.c_int] + [C.c_double] * 7 + [C.c_int, _dp, _ip, _dp, _dp, _ip, _dp, _ip]
            L.ref_solve_mm_O2p.restype = C.c_int
            L.ref_solve_mm_O2p.argtypes = [C.c_double, C.c_double, C.c_int] + [C.c_double] * 9 + [C.c_int, _dp, _ip, _dp, _dp, _ip, _dp, _ip]
            L.ref_ksi_fct2.argtypes = [_dp, C.c_int, _dp, _dp, C.c_int, _dp, _dp, C.c_int, C.c_double, _dp]
            L.ref_spline_eval.argtypes = [_dp, _dp, C.c_int, C.c_int, _dp, C.c_int, _dp]
        if hasattr(L, "ref_Alm"):
            L.ref_Alm.restype = C.c_double
            L.ref_Alm.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
            L.ref_alm_grids_load.restype = C.c_int
            L.ref_alm_grids_load.argtypes = [C.c_char_p]
            L.ref_Alm_interp.restype = C.c_double
            L.ref_Alm_interp.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]

    def Pslm(self, s, l, m):
        return float(self.L.ref_Pslm(s, l, m))

    def Qlm(self, l, m):
        return self.L.ref_Qlm(l, m)

    def amplitude_ratio(self, l, beta):
        V = np.zeros(2 * l + 1)
        self.L.ref_amplitude_ratio(l, float(beta), _p(V))
        return V

    def lin_interpol(self, x, y, xi):
        x, y = _d(x), _d(y)
        return self.L.ref_lin_interpol(_p(x), _p(y), len(x), float(xi))

    def eta0_fct(self, fl0):
        fl0 = _d(fl0)
        return self.L.ref_eta0_fct(_p(fl0), len(fl0))

    def set_imin_imax(self, x, l, fc, gamma, f_s, c, step):
        x = _d(x)
        iv = np.zeros(2, dtype=np.int32)
        self.L.ref_set_imin_imax(_p(x), len(x), l, fc, gamma, f_s, c, step, iv.ctypes.data_as(_ip))
        return int(iv[0]), int(iv[1])

    def build_l_mode_a1etaa3(self, xl, H, fc, fs, eta0, a3, asym, gamma, l, V):
        xl, V = _d(xl), _d(V)
        out = np.zeros_like(xl)
        self.L.ref_build_l_mode_a1etaa3(_p(xl), len(xl), H, fc, fs, eta0, a3, asym, gamma, l, _p(V), _p(out))
        return out

    def build_l_mode_aj(self, xl, H, fc, a, eta0, asym, gamma, l, V):
        xl, V = _d(xl), _d(V)
        out = np.zeros_like(xl)
        self.L.ref_build_l_mode_aj(_p(xl), len(xl), H, fc, *[float(t) for t in a], eta0, asym, gamma, l, _p(V), _p(out))
        return out

    def harvey_like(self, noise, x, y, Nharvey):
        noise, x, y = _d(noise), _d(x), _d(y)
        out = np.zeros_like(x)
        self.L.ref_harvey_like(_p(noise), len(noise), _p(x), _p(y), len(x), Nharvey, _p(out))
        return out

    def chi22p(self, y, model, p=1):
        y, model = _d(y), _d(model)
        return float(self.L.ref_likelihood_chi22p(_p(y), _p(model), len(y), int(p)))

    def chi_square(self, y, model, sigma):
        y, model, sigma = _d(y), _d(model), _d(sigma)
        return float(self.L.ref_likelihood_chi_square(_p(y), _p(model), _p(sigma), len(y)))

    @staticmethod
    def single_thread():
        """One OpenMP thread for the calls that follow: the reference's zeta sums (bump_DP.cpp:137-163) and its model loops
        accumulate under `omp critical`, i.e. in schedule order."""
        C.CDLL("libgomp.so.1").omp_set_num_threads(1)

    def _sols(self, fn, args, cap=4096):
        o = [np.zeros(cap) for _ in range(4)]
        n = [C.c_int(0) for _ in range(3)]
        rc = fn(*args, cap, _p(o[0]), C.byref(n[0]), _p(o[1]), _p(o[2]), C.byref(n[1]), _p(o[3]), C.byref(n[2]))
        assert rc == 0
        return o[0][:n[0].value].copy(), o[1][:n[1].value].copy(), o[2][:n[1].value].copy(), o[3][:n[2].value].copy()

    def solve_mm_from_l0(self, nu_l0, el, delta0l, DPl, alpha, q, resol, fmin, fmax):
        """solve_mm_asymptotic_O2from_l0 (external/ARMM/solver_mm.cpp:624): -> nu_m, nu_p, dnup, nu_g."""
        f = _d(nu_l0)
        return self._sols(self.L.ref_solve_mm_from_l0, (_p(f), len(f), int(el), float(delta0l), float(DPl), float(alpha), float(q), float(resol), float(fmin), float(fmax)))

    def solve_mm_O2p(self, Dnu_p, epsilon, el, delta0l, alpha_p, nmax, DPl, alpha, q, fmin, fmax, resol):
        """solve_mm_asymptotic_O2p (external/ARMM/solver_mm.cpp:470): -> nu_m, nu_p, dnup, nu_g."""
        return self._sols(self.L.ref_solve_mm_O2p, (float(Dnu_p), float(epsilon), int(el), float(delta0l), float(alpha_p), float(nmax), float(DPl), float(alpha),
                                                    float(q), float(fmin), float(fmax), float(resol)))

    def spline_eval(self, x, y, xq, kind=1):
        x, y, xq = _d(x), _d(y), _d(xq)
        o = np.zeros(len(xq))
        self.L.ref_spline_eval(_p(x), _p(y), len(x), int(kind), _p(xq), len(xq), _p(o))
        return o

    def Alm(self, l, m, theta0, delta, filter_code=0):
        """The reference's direct integral Alm() (activity.cpp:221-246); radians."""
        return self.L.ref_Alm(int(l), int(m), float(theta0), float(delta), int(filter_code))

    def alm_grids_load(self, grid_dir):
        """Config::Config's grid set-up (config.cpp:77-147) with the reference's own loadAllData / init_2dgrid."""
        return self.L.ref_alm_grids_load(os.fsencode(grid_dir))

    def Alm_interp(self, l, m, theta0, delta, filter_code=0):
        """Alm_interp_iter_preinitialised (Alm_interpol.cpp:188-348) on the grids loaded by alm_grids_load."""
        return self.L.ref_Alm_interp(int(l), int(m), float(theta0), float(delta), int(filter_code))

    def call_model(self, model_id, params, plength, x):
        params, x = _d(params), _d(x)
        pl = np.ascontiguousarray(plength, dtype=np.int32)
        out = np.zeros(len(x))
        rc = self.L.ref_call_model(model_id, _p(params), len(params), pl.ctypes.data_as(_ip), _p(x), len(x), _p(out))
        return rc, out


    def call_model_recorded(self, model_id, params, plength, x, cap=4096):
        """The reference's model function + one row {l, fc, H, W, a1..a6, eta0, asym, step, c, V[7]} per
        optimum_lorentzian_calc_aj call it made (the mode list its host code resolved, e.g. the ARMM mixed modes)."""
        params, x = _d(params), _d(x)
        pl = np.ascontiguousarray(plength, dtype=np.int32)
        out = np.zeros(len(x))
        rows = np.zeros((cap, 21))
        n = C.c_int(0)
        rc = self.L.ref_call_model_recorded(model_id, _p(params), len(params), pl.ctypes.data_as(_ip), _p(x), len(x), _p(out),
                                            _p(rows), cap, C.byref(n))
        assert n.value <= cap
        return rc, out, rows[: n.value].copy()

    def eval_chains(self, model_id, params, plength, x, y, Tcoefs, p=1.0, nthreads=0):
        """The reference's per-chain OpenMP fan-out of call_model + call_likelihood (MALA.cpp:648, model_def.cpp:466-482)."""
        params = _d(params)
        Nchains, Nparams = params.shape
        x, y, T = _d(x), _d(y), _d(Tcoefs)
        pl = np.ascontiguousarray(plength, dtype=np.int32)
        out = np.zeros(Nchains)
        rc = self.L.ref_eval_chains(model_id, _p(params), Nparams, pl.ctypes.data_as(_ip), _p(x), _p(y), len(x), Nchains, _p(T),
                                    float(p), _p(out), int(nthreads))
        return rc, out

    def max_threads(self):
        return int(self.L.ref_max_threads())


_inst = None
_inst_O3 = None


def get_O3():
    global _inst_O3
    if _inst_O3 is None:
        _inst_O3 = Ref(SO_O3)
    return _inst_O3


def get():
    global _inst
    if _inst is None:
        _inst = Ref()
    return _inst
