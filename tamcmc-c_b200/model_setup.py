"""From a parsed MS_Global `.model` file to the fit's parameter vector: the host-side set-up step in front of the hot path.

Restates `build_init_MS_Global` (tamcmc/sources/io_ms_global.cpp:362-1400) with its helpers `set_noise_params` (:1402-1536),
`getnumax` (:1559-1574), `settings_aj_splittings` (:1718-1870) and the `IO_models` block operations
(tamcmc/sources/io_models.cpp: `initialise_param` :244-281, `fill_param` :44-77, `add_param` :120-141): the fields of `MCMC_files`
(formats.read_ms_global_model) become

    inputs (the flat parameter vector in the layout the model functions unpack, io_ms_global.cpp:1315-1398),
    plength (11 block lengths: heights, visibilities, l=0..3 frequencies, splittings, widths, noise, inclination, switches),
    relax (which entries the sampler moves), priors_names / priors (4 x N prior table), inputs_names, extra_priors (10 values).

`inputs` and `plength` are what `tamcmc_gpu_create` / `tamcmc_gpu_eval` take (include/tamcmc_gpu.h); `relax` and the prior table
feed the driver (host/mcmc_driver.hpp, host/priors.hpp).  Covered: every model name the function accepts except the two
Appourchaux-width variants (`model_MS_Global_a1etaa3_AppWidth_HarveyLike_v1/_v2`, which no GPU model id serves).  Where the
reference prints a message and calls exit(), this raises ValueError with the same diagnosis.  `build_init_asymptotic` (io_asymptotic.cpp:32-875)
is the red-giant dialect, `build_init_local` (io_local.cpp:329-1238) the local-fit one (models 11 / 14; formats.read_local_model).  New code: the reference's blocks of
`if (name == ...)` are table-driven here; pinned value for value on the reference's own function, compiled from its own sources
(tests/test_model_setup.py, tests/golden/reference_ms_global_init.json)."""
import math

import numpy as np

NPRIOR = 4          # Nmax_prior_params (io_ms_global.cpp:373)
LD = np.longdouble
PI_LD = LD("3.141592653589793238")          # `const long double pi` (io_ms_global.cpp:364): expressions that contain it run in extended precision
EMPTY = -9999.0

# model_fullname -> (do_a11_eq_a12, do_avg_a1n, aj_switch, extra_priors[9])   io_ms_global.cpp:430-509
_MODEL_RULES = {
    "model_MS_Global_a1etaa3_HarveyLike_Classic": (1, 1, 0, None),
    "model_MS_Global_a1etaa3_HarveyLike_Classic_v2": (1, 1, 0, None),
    "model_MS_Global_a1etaa3_HarveyLike_Classic_v3": (1, 1, 0, None),
    "model_MS_Global_a1etaa3_HarveyLike": (1, 1, 0, 0),
    "model_MS_Global_a1etaa3_Harvey1985": (1, 1, 0, 0),
    "model_MS_Global_a1a2a3_HarveyLike": (1, 1, 1, 1),
    "model_MS_Global_a1l_etaa3_HarveyLike": (0, 1, 0, 2),
    "model_MS_Global_a1n_etaa3_HarveyLike": (1, 0, 0, 3),
    "model_MS_Global_a1nl_etaa3_HarveyLike": (0, 0, 0, 4),
    "model_MS_Global_a1n_a2a3_HarveyLike": (1, 0, 2, 5),
    "model_MS_Global_a1l_a2a3_HarveyLike": (0, 1, 3, 6),
    "model_MS_Global_a1nl_a2a3_HarveyLike": (0, 0, 4, 7),
    "model_MS_Global_ajAlm_HarveyLike": (1, 1, 5, 8),
    "model_MS_Global_aj_HarveyLike": (1, 1, 6, 9),
}
_SQRT_A1_MODELS = ("model_MS_Global_a1etaa3_HarveyLike", "model_MS_Global_a1etaa3_Harvey1985", "model_MS_Global_a1etaa3_AppWidth_HarveyLike_v1",
                   "model_MS_Global_a1etaa3_AppWidth_HarveyLike_v2", "model_MS_Global_a1n_a2a3_HarveyLike", " model_MS_Global_a1nl_a2a3_HarveyLike",
                   "model_MS_Global_a1a2a3_HarveyLike")          # (the stray blank in the sixth name is the reference's, :1195)
# keyword -> position in the splitting block: aj model (aj_switch 6), ajAlm model (aj_switch 5)   io_ms_global.cpp:1718-1870
_AJ_POS6 = {"a1_0": 0, "a1_1": 1, "a2_0": 2, "a2_1": 3, "a3_0": 4, "a3_1": 5, "a4_0": 6, "a4_1": 7, "a5_0": 8, "a5_1": 9, "a6_0": 10, "a6_1": 11}
_AJ_POS5 = {"a1_0": 0, "a1_1": 1, "a3_0": 2, "a3_1": 3, "a5_0": 4, "a5_1": 5}
_EPS_POS5 = {"epsilon_0": 6, "epsilon_1": 7, "theta0": 8, "delta": 9}


class Block:
    """A block of parameters: IO_models::initialise_param (io_models.cpp:244-281)."""

    def __init__(self, n):
        self.names = ["Empty"] * n
        self.pnames = ["Fix"] * n
        self.inputs = np.zeros(n)
        self.relax = np.zeros(n, dtype=np.int64)
        self.priors = np.full((NPRIOR, n), EMPTY)

    def fill(self, name, prior, val, prior_vals, pos, i0):
        """IO_models::fill_param (io_models.cpp:44-77)"""
        self.names[pos] = name
        self.pnames[pos] = prior
        self.inputs[pos] = val
        if prior == "Fix":
            self.relax[pos] = 0
            self.priors[:, pos] = EMPTY
        else:
            self.relax[pos] = 1
            self.priors[:, pos] = [prior_vals[k + i0] for k in range(NPRIOR)]

    def __len__(self):
        return len(self.names)


def _fatal(var, kind):
    raise ValueError("%s: %s (fatalerror_msg_io_MS_Global, io_ms_global.cpp:1538-1557)" %
                     (var, "Fix_Auto is not implemented for that parameter" if kind == "Fix_Auto" else "should always be defined as '%s'" % kind))


def lin_interpol(x, y, x_int):
    """tamcmc/sources/interpol.cpp:13-43"""
    n = len(x)
    i, a, b = 0, 0.0, 0.0
    if x[0] <= x_int <= x[n - 1]:
        while (x_int < x[i] or x_int > x[i + 1]) and i < n - 2:
            i += 1
        a = (y[i + 1] - y[i]) / (x[i + 1] - x[i])
        b = y[i] - a * x[i]
    if x_int < x[0]:
        a = (y[1] - y[0]) / (x[1] - x[0])
        b = y[0] - a * x[0]
    if x_int > x[n - 1]:
        a = (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2])
        b = y[n - 2] - a * x[n - 2]
    return a * x_int + b


def set_noise_params(noise_s2, noise_params):
    """io_ms_global.cpp:1402-1536: three Harvey profiles + white noise; the high-frequency profile and the white noise carry
    Gaussian priors centred on the file's values."""
    nz = Block(10)
    nz.names = ["Harvey-Noise_H", "Harvey-Noise_tc", "Harvey-Noise_p"] * 3 + ["White_Noise_N0"]
    nz.pnames = ["Fix"] * 6 + ["Gaussian"] * 4
    nz.relax[:] = [0] * 6 + [1] * 4
    nz.inputs = np.array(noise_params, dtype=np.float64).copy()
    for k in (0, 3, 6):
        if nz.inputs[k] <= 0 or nz.inputs[k + 1] <= 0 or nz.inputs[k + 2] <= 0:
            nz.pnames[k:k + 3] = ["Fix"] * 3
            nz.relax[k:k + 3] = 0
            nz.inputs[k:k + 3] = [0.0, 0.0, 1.0]
    for k in (6, 7, 8, 9):
        nz.priors[0, k] = noise_s2[k, 0]
    nz.priors[1, 6] = (noise_s2[6, 1] + noise_s2[6, 2]) * 3. / 2
    nz.priors[1, 7] = (noise_s2[7, 1] + noise_s2[7, 2]) * 3. / 2
    nz.priors[1, 8] = (noise_s2[8, 1] + noise_s2[8, 2]) * 3. / 2 if noise_s2[8, 1] != 0 else nz.priors[0, 8] * 0.1
    nz.priors[1, 9] = (noise_s2[9, 1] + noise_s2[9, 2]) if nz.pnames[9] == "Uniform" else nz.priors[0, 9] * 0.1
    with np.errstate(divide="ignore", invalid="ignore"):
        for k, floor in ((6, 0.05), (7, 0.005), (8, 0.05), (9, 0.0005)):
            if nz.priors[1, k] / nz.priors[0, k] <= floor and nz.pnames[k] != "Fix":
                nz.priors[1, k] = nz.priors[0, k] * floor
    return nz


def amplitude_ratio(l, beta_deg):
    """tamcmc/sources/function_rot.cpp:15-101 (host copy; the device expander has its own, expand.cu)."""
    def combi(n, r):
        return (math.factorial(n) // math.factorial(r)) // math.factorial(n - r)      # integer divisions like function_rot.cpp:90-92

    def dmm(l, m1, m2, beta):
        var = 0.0
        if m1 + m2 >= 0:
            for s in range(0, l - m1 + 1):
                if l - m1 - s >= 0 and l + m2 >= l - m1 - s and l - m2 >= s:
                    var += combi(l + m2, l - m1 - s) * combi(l - m2, s) * (-1.0) ** (l - m1 - s) * math.cos(beta / 2.) ** (2 * s + m1 + m2) * math.sin(beta / 2.) ** (2 * l - 2 * s - m1 - m2)
            return var * math.sqrt(math.factorial(l + m1) * math.factorial(l - m1)) / math.sqrt(math.factorial(l + m2) * math.factorial(l - m2))
        return None

    angle = math.pi * beta_deg / 180.
    out = np.zeros(2 * l + 1)
    for m in range(-l, l + 1):
        a = abs(m)
        d = dmm(l, a, 0, angle)
        out[m + l] = d * d
    return out


def build_init_ms_global(mf, resol):
    """mf: the dictionary of formats.read_ms_global_model; resol: the spectrum's resolution (Width lower bound of Fix_Auto).
    Returns a dictionary with model_fullname, inputs, relax, priors (4 x N), plength (11), extra_priors (10), inputs_names, priors_names."""
    Hmin, Hmax = 1.0, 10000.0
    Vl = [1, 1.5, 0.53, 0.08]
    Dnu = mf["Dnu"]
    numax, err_numax = mf["numax"], mf["err_numax"]
    names, cpri, mc = mf["common_names"], mf["common_names_priors"], mf["modes_common"]
    extra = np.array([1, 2., 1e6, 0.50, 0.20, 0.15, 0.05, 0.05, 0, -1], dtype=np.float64)      # :406-417
    els = np.asarray(mf["els"])
    lmax = int(els.max())

    # ---- instructions that come before the set-up (:424-543) ----
    fullname, do_amp, filter_type = " ", 0, ""
    do_a11_eq_a12, do_avg_a1n, aj_switch = 1, 1, 0
    for i, nm in enumerate(names):
        if nm == "model_fullname":
            fullname = cpri[i]
            if fullname in ("model_MS_Global_a1etaa3_AppWidth_HarveyLike_v1", "model_MS_Global_a1etaa3_AppWidth_HarveyLike_v2"):
                raise ValueError("%s: the Appourchaux-width MS models are not restated here (no GPU model id serves them)" % fullname)
            if fullname in _MODEL_RULES:
                do_a11_eq_a12, do_avg_a1n, sw, e9 = _MODEL_RULES[fullname]
                aj_switch = sw
                if e9 is not None:
                    extra[9] = e9
        if nm == "fit_squareAmplitude_instead_Height":
            if cpri[i] != "bool":
                _fatal(nm, "bool")
            do_amp = int(mc[i, 0])
        if nm == "filter_type":
            filter_type = cpri[i]
    if fullname == " ":
        raise ValueError("Model name empty: the .model file needs the model_fullname variable (io_ms_global.cpp:544-547)")
    vis, inc = Block(lmax), Block(1)

    # ---- frequencies / widths / heights of the eigen table, degree by degree, matched with the relax list (:555-612) ----
    eig = np.asarray(mf["eigen_params"], dtype=np.float64)
    f_inputs, f_min, f_max, w_inputs, h_inputs, f_relax, w_relax, h_relax = [], [], [], [], [], [], [], []
    Nf_el = [0, 0, 0, 0]
    for el in range(lmax + 1):
        pos0 = [k for k in range(len(els)) if els[k] == el]
        f_el = [mf["freqs_ref"][k] for k in pos0]
        pos_el = [k for k in range(eig.shape[0]) if int(eig[k, 0]) == el]
        Nf_el[el] = len(pos_el)
        for k in pos_el:
            f_inputs.append(eig[k, 1]); f_min.append(eig[k, 2]); f_max.append(eig[k, 3])
            if el == 0:
                w_inputs.append(eig[k, 4]); h_inputs.append(eig[k, 5])
            hits = [j for j in range(len(f_el)) if eig[k, 1] - 1e-2 <= f_el[j] <= eig[k, 1] + 1e-2]      # where_dbl, string_handler.cpp:88-110
            if len(hits) != 1:
                raise ValueError("the frequency %r is not unique in / absent from the relax list (io_ms_global.cpp:586-601)" % eig[k, 1])
            f_relax.append(bool(mf["relax_freq"][pos0[hits[0]]]))
            if el == 0:
                w_relax.append(bool(mf["relax_gamma"][pos0[hits[0]]])); h_relax.append(bool(mf["relax_H"][pos0[hits[0]]]))

    # ---- defaults (:617-690) ----
    if do_amp:
        h_name = "Amplitude_l0"
        h_inputs = [float(PI_LD * LD(w_inputs[k]) * LD(h_inputs[k])) for k in range(len(h_inputs))]
    else:
        h_name = "Height_l0"
    height, width, freq = Block(len(h_relax)), Block(len(w_relax)), Block(len(f_relax))
    tmp = [Hmin, Hmax, EMPTY, EMPTY]
    for k in range(len(h_inputs)):
        height.fill(h_name, "Jeffreys" if h_relax[k] else "Fix", h_inputs[k], tmp, k, 0)
    tmp = [resol, Dnu / 3., EMPTY, EMPTY]
    for k in range(len(w_inputs)):
        width.fill("Width_l0", "Jeffreys" if w_relax[k] else "Fix", w_inputs[k] if w_inputs[k] < Dnu / 3. else Dnu / 3.1, tmp, k, 0)
    for k in range(len(f_inputs)):
        tmp = [f_min[k], f_max[k], 0.01 * Dnu, 0.01 * Dnu]
        freq.fill("Frequency_l", "GUG" if f_relax[k] else "Fix", f_inputs[k], tmp, k, 0)
    # ---- numax (:692-723) ----
    if numax <= 0:
        Hflat = np.zeros(sum(Nf_el))
        Hflat[:Nf_el[0]] = height.inputs
        cpt = Nf_el[0]
        for el in range(1, 4):
            for n in range(Nf_el[el]):
                # (the reference interpolates at freq_in.inputs[Nf_el[0] + n] for every degree, :705)
                Hflat[cpt + n] = abs(lin_interpol(freq.inputs[:Nf_el[0]], height.inputs, freq.inputs[Nf_el[0] + n])) * Vl[el]
            cpt += Nf_el[el]
        numax = _getnumax(freq.inputs, Hflat)
    elif err_numax <= 0:
        err_numax = 0.05 * numax

    # ---- size of the splitting block (:729-838) ----
    n1, n2, n0 = Nf_el[1], Nf_el[2], Nf_el[0]
    if do_a11_eq_a12 == 1 and do_avg_a1n == 1:
        sizes = {0: 6, 1: 9, 2: 6 + n0, 3: 6 + lmax * 3, 4: 6 + n0, 5: 12, 6: 14}
    elif do_a11_eq_a12 == 0 and do_avg_a1n == 1:
        sizes = {0: 7, 1: 10, 2: 7 + n0, 3: 7 + lmax * 3, 4: 7 + n0}
    elif do_a11_eq_a12 == 1 and do_avg_a1n == 0:
        if n1 != n2:
            raise ValueError("a11 = a22 needs as many l=1 as l=2 modes: %d and %d (io_ms_global.cpp:808-815)" % (n1, n2))
        sizes = {0: 6 + n1, 1: 6 + n1 + 3, 2: 6 + n1 + n0, 3: 6 + n1 + lmax * 3, 4: 6 + n1 + n0}
    else:
        sizes = {0: 6 + n1 + n2, 1: 6 + n1 + n2 + 3, 2: 6 + n1 + n2 + n0, 3: 6 + n1 + n2 + lmax * 3, 4: 6 + n1 + n2 + n0}
    if aj_switch not in sizes:
        raise ValueError("aj_switch = %d has no preset rule for this model family (io_ms_global.cpp:759-764)" % aj_switch)
    snlm = Block(sizes[aj_switch])
    if aj_switch == 5:
        snlm.fill("eta0_switch", "Fix", 1, [EMPTY] * 4, 10, 0)
    if aj_switch == 6:
        snlm.fill("eta0_switch", "Fix", 0, [EMPTY] * 4, 12, 0)

    # ---- the common parameters, in file order (:840-1166) ----
    trunc_c, decompose_Alm = -1.0, -1
    a2_count, aj_count = 0, 0
    bool_a1sini = bool_a1cosi = False
    for i, nm in enumerate(names):
        pr, row = cpri[i], mc[i]
        if nm in ("freq_smoothness", "Freq_smoothness"):
            if pr != "bool":
                _fatal("freq_smoothness", "bool")
            extra[0], extra[1] = row[0], row[1]
        if nm == "trunc_c":
            if pr != "Fix":
                try:
                    trunc_c = float(pr)
                except ValueError:
                    _fatal("trunc_c", "Fix")
            else:
                trunc_c = row[0]
        if nm in ("Frequency", "frequency"):
            if pr not in ("GUG", "Uniform"):
                _fatal(nm, "GUG or Uniform")
            for k in range(len(f_inputs)):
                tmp = [f_min[k], f_max[k], row[3], row[4]] if pr == "GUG" else [f_min[k], f_max[k], EMPTY, EMPTY]
                freq.fill("Frequency_l", pr if f_relax[k] else "Fix", f_inputs[k], tmp, k, 0)
        if nm in ("height", "Height", "amplitude", "Amplitude"):
            if pr == "Fix_Auto":
                _fatal(nm, "Fix_Auto")
            for k in range(len(h_inputs)):
                if h_relax[k]:
                    height.fill(h_name, pr, h_inputs[k], row, k, 0)
                else:
                    height.fill(h_name, "Fix", h_inputs[k], row, k, 1)
        if nm in ("width", "Width"):
            if pr == "Fix_Auto":
                wprior, wvals = "Jeffreys", [resol, Dnu / 3., EMPTY, EMPTY]
            else:
                wprior, wvals = pr, row
            for k in range(len(w_inputs)):
                val = w_inputs[k] if w_inputs[k] < Dnu / 3. else Dnu / 3.1
                if w_relax[k]:
                    width.fill("Width_l", wprior, val, wvals, k, 0)
                else:
                    width.fill("Width_l", "Fix", val, row, k, 1)
        if nm in ("splitting_a1", "Splitting_a1"):
            if pr == "Fix_Auto":
                _fatal("splitting_a1", "Fix_Auto")
            snlm.fill("Splitting_a1", pr, row[0], row, 0, 1)
            if do_a11_eq_a12 == 0 and do_avg_a1n == 1:
                snlm.fill(snlm.names[0], snlm.pnames[0], snlm.inputs[0], row, 6, 1)
            if do_a11_eq_a12 == 1 and do_avg_a1n == 0:
                for kk in range(n1):
                    snlm.fill(snlm.names[0], snlm.pnames[0], snlm.inputs[0], row, 6 + kk, 1)
                snlm.fill("Empty", "Fix", 0, row, 0, 1)
            if do_a11_eq_a12 == 0 and do_avg_a1n == 0:
                for kk in range(n1 + n2):
                    snlm.fill(snlm.names[0], snlm.pnames[0], snlm.inputs[0], row, 6 + kk, 1)
                snlm.fill("Empty", "Fix", 0, row, 0, 1)
        if nm in ("asphericity_eta", "Asphericity_eta"):                 # ignored since 07/12/2021 (:951-960)
            snlm.pnames[1] = "Fix"; snlm.names[1] = "Asphericity_eta"; snlm.relax[1] = 0; snlm.inputs[1] = 0
        if nm == "a2":
            raise ValueError("the a2 keyword is not usable (io_ms_global.cpp:961-972): the a1a2a3 model takes a2_0, a2_1, a2_2")
        if nm in ("a2_0", "a2_1", "a2_2") and aj_switch == 1:
            a2_count += 1
            if pr == "Fix_Auto":
                raise ValueError("Fix_Auto requested for %s: not allowed" % nm)
            snlm.fill(nm, pr, row[0], row, 6 + int(nm[-1]), 1)
        pos_tab = _AJ_POS6 if aj_switch == 6 else _AJ_POS5 if aj_switch == 5 else {}
        if nm in pos_tab:
            aj_count += 1
            if pr == "Fix_Auto":
                raise ValueError("Fix_Auto requested for %s: not allowed" % nm)
            snlm.fill(nm, pr, row[0], row, pos_tab[nm], 1)
        if nm in _EPS_POS5 and aj_switch == 5:
            a2_count += 1
            if pr == "Fix_Auto":
                raise ValueError("Fix_Auto requested for %s: not allowed" % nm)
            snlm.fill(nm, pr, row[0], row, _EPS_POS5[nm], 1)
        if nm == "decompose_Alm" and aj_switch == 5:
            if pr != "Fix":
                _fatal("decompose_Alm", "Fix")
            decompose_Alm = int(row[0])
        if nm in ("splitting_a3", "Splitting_a3"):
            if pr == "Fix_Auto":
                _fatal("splitting_a3", "Fix_Auto")
            snlm.fill("Splitting_a3", pr, row[0], row, 2, 1)
        if nm in ("asymetry", "Asymetry"):
            if pr == "Fix_Auto":
                _fatal("asymetry", "Fix_Auto")
            snlm.fill("Lorentzian_asymetry", pr, row[0], row, 5 if aj_switch not in (5, 6) else len(snlm) - 1, 1)
        for l in (1, 2, 3):
            if nm in ("visibility_l%d" % l, "Visibility_l%d" % l):
                if pr == "Fix_Auto":
                    _fatal("visibility_l%d" % l, "Fix_Auto")
                if lmax >= l:
                    vis.fill("Visibility_l%d" % l, pr, row[0], row, l - 1, 1)
        if nm in ("inclination", "Inclination"):
            if pr == "Fix_Auto":
                _fatal("inclination", "Fix_Auto")
            inc.fill("Inclination", pr, 89.99999 if row[0] >= 90 else row[0], row, 0, 1)
        if nm in ("sqrt(splitting_a1).cosi", "sqrt(splitting_a1).sini"):
            if pr == "Fix_Auto":
                _fatal(nm, "Fix_Auto")
            snlm.fill(nm, pr, row[0], row, 3 if nm.endswith("cosi") else 4, 1)
            if do_a11_eq_a12 == 0 or do_avg_a1n == 0:
                raise ValueError("sqrt(a1).cosi / sqrt(a1).sini are not available for the a1n / a1l models (io_ms_global.cpp:1143-1148)")
            if nm.endswith("cosi"):
                bool_a1cosi = True
            else:
                bool_a1sini = True
    if aj_switch == 1 and a2_count != 3:
        raise ValueError("Invalid number of constraints for a2: set a2_0, a2_1 and a2_2")
    if aj_switch == 5 and a2_count != 4:
        raise ValueError("Invalid number of constraints: set epsilon_0, epsilon_1, theta0 and delta")
    if aj_switch == 6 and aj_count != 12:
        raise ValueError("Invalid number of constraints: set a1_0 ... a6_1 (12 parameters)")
    if bool_a1cosi != bool_a1sini:
        raise ValueError("both sqrt(splitting_a1).sini and sqrt(splitting_a1).cosi must appear (io_ms_global.cpp:1180-1186)")

    if not bool_a1cosi and not bool_a1sini:
        if fullname in _SQRT_A1_MODELS:
            # splitting_a1 and inclination are replaced by sqrt(a1) cos i and sqrt(a1) sin i (:1187-1226)
            col0 = snlm.priors[:, 0].copy()
            if inc.pnames[0] == "Fix" and snlm.pnames[0] == "Fix":
                snlm.fill("sqrt(splitting_a1).cosi", "Fix", _proj(snlm.inputs[0], inc.inputs[0], np.cos), col0, 3, 0)
                snlm.fill("sqrt(splitting_a1).sini", "Fix", _proj(snlm.inputs[0], inc.inputs[0], np.sin), col0, 4, 0)
            else:
                snlm.priors[1, 0] = math.sqrt(snlm.priors[1, 0])
                col0 = snlm.priors[:, 0].copy()
                snlm.fill("sqrt(splitting_a1).cosi", snlm.pnames[0], _proj(snlm.inputs[0], inc.inputs[0], np.cos), col0, 3, 0)
                snlm.fill("sqrt(splitting_a1).sini", snlm.pnames[0], _proj(snlm.inputs[0], inc.inputs[0], np.sin), col0, 4, 0)
            if snlm.inputs[3] < 1e-2:
                snlm.inputs[3] = 1e-2
            if snlm.inputs[4] < 1e-2:
                snlm.inputs[4] = 1e-2
            inc.fill("Empty", "Fix", 0, inc.priors[:, 0].copy(), 0, 1)
            snlm.fill("Empty", "Fix", 0, snlm.priors[:, 0].copy(), 0, 1)
        if fullname == "model_MS_Global_a1etaa3_HarveyLike_Classic_v2":
            # the inclination block becomes the m-height ratios: 2 for l=1, 3 for l=2, 4 for l=3 (:1234-1248)
            inc0 = inc.inputs[0]
            inc = Block(9)
            ind = 0
            for el in range(1, lmax + 1):
                r = amplitude_ratio(el, inc0)
                for em in range(el + 1):
                    inc.fill("Inc:H%d,%d" % (el, em), "Uniform", r[el + em], [0, 1, EMPTY, EMPTY], ind, 0)
                    ind += 1
            extra[8] = 1
        if fullname == "model_MS_Global_a1etaa3_HarveyLike_Classic_v3":
            # one height per (n, l, |m|); the visibilities are switched off (:1249-1281)
            inc0 = inc.inputs[0]
            vis_vals = vis.inputs.copy()
            for el in range(1, lmax):
                vis.fill("Empty", "Fix", 0, [EMPTY] * 4, el - 1, 0)
            inc = Block(Nf_el[1] * 2 + Nf_el[2] * 3 + Nf_el[3] * 4)
            ind = 0
            for el in range(1, lmax + 1):
                r = amplitude_ratio(el, inc0)
                for en in range(Nf_el[el]):
                    for em in range(el + 1):
                        inc.fill("Inc: H%d,%d,%d" % (en, el, em), "Jeffreys", height.inputs[en] * vis_vals[el - 1] * r[el + em], [Hmin, Hmax, EMPTY, EMPTY], ind, 0)
                        ind += 1
            extra[8] = 2
    else:
        if fullname == "model_MS_Global_a1etaa3_HarveyLike_Classic":
            raise ValueError("%s cannot be used with sqrt(splitting_a1).cosi / .sini (io_ms_global.cpp:1289-1295)" % fullname)
        inc.fill("Empty", "Fix", 0, inc.priors[:, 0].copy(), 0, 1)
        snlm.fill("Empty", "Fix", 0, snlm.priors[:, 0].copy(), 0, 1)

    noise = set_noise_params(np.asarray(mf["noise_s2"], dtype=np.float64), mf["noise_params"])

    # ---- everything in one vector (:1311-1398) ----
    alm = fullname == "model_MS_Global_ajAlm_HarveyLike"
    plength = np.array([len(h_inputs), lmax, Nf_el[0], Nf_el[1], Nf_el[2], Nf_el[3], len(snlm), len(w_inputs), len(noise), len(inc), 4 if alm else 2], dtype=np.int64)
    allp = Block(int(plength.sum()))
    p0 = 0
    for blk in (height, vis, freq, snlm, width, noise, inc):
        n = len(blk)
        allp.names[p0:p0 + n] = blk.names
        allp.pnames[p0:p0 + n] = blk.pnames
        allp.inputs[p0:p0 + n] = blk.inputs
        allp.relax[p0:p0 + n] = blk.relax
        allp.priors[:, p0:p0 + n] = blk.priors
        p0 += n
    row0 = mc[0]
    allp.fill("Truncation parameter", "Fix", trunc_c, row0, p0, 1)
    if allp.inputs[p0] <= 0:
        allp.inputs[p0] = 10000.
    allp.fill("Switch for fit of Amplitudes or Heights", "Fix", do_amp, row0, p0 + 1, 1)
    if alm:
        allp.fill("decompose_Alm", "Fix", decompose_Alm, row0, p0 + 2, 1)
        codes = {"gate": 0, "gauss": 1, "triangle": 2}
        if filter_type not in codes:
            raise ValueError("Unrecognized filter type %r: gate, gauss or triangle (io_ms_global.cpp:1383-1386)" % filter_type)
        allp.fill("filter_type", "Fix", codes[filter_type], [EMPTY] * 4, p0 + 3, 0)
    if aj_switch >= 2 and aj_switch not in (5, 6):
        raise ValueError("aj_switch >= 2 is not configured in the reference (io_ms_global.cpp:1390-1394: it exits)")
    return {"model_fullname": fullname, "inputs": allp.inputs, "relax": allp.relax, "priors": allp.priors, "plength": plength,
            "extra_priors": extra, "inputs_names": allp.names, "priors_names": allp.pnames, "numax": numax, "err_numax": err_numax}


_RGB_SPLIT_POS = {"rot_env": 0, "Rot_env": 0, "a1_env": 0, "rot_core": 1, "Rot_core": 1, "a1_core": 1, "a2_core": 2, "a2_env": 3,
                  "a3_env": 4, "a4_env": 5, "a5_env": 6, "a6_env": 7}          # settings_aj_splittings_RGB, io_asymptotic.cpp:877-984
_RGB_SPLIT_NAME = {0: "rot_env", 1: "rot_core", 2: "a2_core", 3: "a2_env", 4: "a3_env", 5: "a4_env", 6: "a5_env", 7: "a6_env"}
_RGB_L1_POS = {"DP1": 1, "alpha_g": 2, "q": 3, "sigma_Hl1": 4, "Wfactor": 6, "Hfactor": 7}      # global parameters of the l=1 mixed modes (:513-560)
RGB_V4_MODELS = {"model_RGB_asympt_aj_AppWidth_HarveyLike_v4": 25, "model_RGB_asympt_aj_CteWidth_HarveyLike_v4": 27}      # -> tamcmc_host_expand_rgb_v4 ids


def set_width_app2016_params_v2(numax, err_numax):
    """io_ms_global.cpp:1625-1716: initial guesses and priors of the six parameters of the Appourchaux et al. 2016 width relation"""
    w = Block(6)
    out = [abs(numax), abs(numax), abs(4. / 2150. * numax + (1. - 1000. * 4. / 2150.)), abs(0.8 / 2150. * numax + (4.5 - 1000. * 0.8 / 2150.)),
           abs(3400. / 2150. * numax + (1000. - 1000. * 3400. / 2150.)), abs(2.8 / 2200. * numax + (1. - 2.8 / 2200. * 1.))]
    if numax < 800:
        out[3] = out[3] / 5
    pri = [[out[0], err_numax], [out[1], err_numax], [0, 6], [0, 10], [out[4], out[4] * 0.25], [0., 15]]
    kinds = ["Gaussian", "Gaussian", "Uniform", "Uniform", "Gaussian", "Uniform"]
    for k, nm in enumerate(("numax", "nudip", "alpha", "Gamma_alpha", "Wdip", "DeltaGammadip")):
        w.fill("width:Appourchaux_v2:" + nm, kinds[k], out[k], pri[k] + [EMPTY, EMPTY], k, 0)
    return w


def build_init_asymptotic(mf, resol):
    """The red-giant dialect: build_init_asymptotic (tamcmc/sources/io_asymptotic.cpp:32-875) for the two models it still
    accepts, model_RGB_asympt_aj_{AppWidth,CteWidth}_HarveyLike_v4 (every other name makes the reference exit, :86-92, :139-141).
    The l=1 block of the frequency section holds the global parameters of the mixed modes (delta01, DP1, alpha_g, q, sigma_Hl1, -,
    Wfactor, Hfactor) followed by the nodes (fref) and values (ferr) of the bias spline taken from the hyper-prior rows; the result
    is the vector tamcmc_host_expand_rgb_v4 (csrc/host_rgb.cpp) turns into a mode-table row."""
    Hmin, Hmax = 1.0, 10000.0
    NG = 7                                          # Nmixedmodes_g_params
    Dnu = mf["Dnu"]
    sigma_limit = Dnu / 10.
    numax, err_numax = mf["numax"], mf["err_numax"]
    names, cpri, mc = mf["common_names"], mf["common_names_priors"], mf["modes_common"]
    els = np.asarray(mf["els"])
    lmax = int(els.max())
    fullname, do_amp, dwa = " ", 0, 0
    for i, nm in enumerate(names):
        if nm == "model_fullname":
            fullname = cpri[i]
            if fullname == "model_RGB_asympt_aj_AppWidth_HarveyLike_v4":
                dwa = 2
                if numax <= 0:
                    raise ValueError("%s needs a positive numax (!n line) (io_asymptotic.cpp:97-106)" % fullname)
            elif fullname == "model_RGB_asympt_aj_CteWidth_HarveyLike_v4":
                dwa = 1
                if numax != -9999 and numax <= 0:
                    raise ValueError("%s: numax must be positive or absent (io_asymptotic.cpp:113-122)" % fullname)
        if nm == "fit_squareAmplitude_instead_Height":
            if cpri[i] != "bool":
                _fatal(nm, "bool")
            do_amp = int(mc[i, 0])
    if fullname not in RGB_V4_MODELS:
        raise ValueError("model name %r: only the two *_aj_*Width_HarveyLike_v4 models are accepted (io_asymptotic.cpp:135-141)" % fullname)
    vis, inc = Block(lmax), Block(1)

    eig = np.asarray(mf["eigen_params"], dtype=np.float64)
    f_inputs, f_min, f_max, w_inputs, h_inputs, f_relax, w_relax, h_relax = [], [], [], [], [], [], [], []
    Nf_el = [0, 0, 0, 0]
    for el in range(lmax + 1):
        pos0 = [k for k in range(len(els)) if els[k] == el]
        f_el = [mf["freqs_ref"][k] for k in pos0]
        pos_el = [k for k in range(eig.shape[0]) if int(eig[k, 0]) == el]
        Nf_el[el] = len(pos_el)
        for k in pos_el:
            f_inputs.append(eig[k, 1]); f_min.append(eig[k, 2]); f_max.append(eig[k, 3])
            if el == 0:
                w_inputs.append(eig[k, 4]); h_inputs.append(eig[k, 5])
            hits = [j for j in range(len(f_el)) if eig[k, 1] - 1e-2 <= f_el[j] <= eig[k, 1] + 1e-2]
            if len(hits) != 1:
                raise ValueError("the frequency %r is not unique in / absent from the relax list (io_asymptotic.cpp:176-188)" % eig[k, 1])
            f_relax.append(bool(mf["relax_freq"][pos0[hits[0]]]))
            if el == 0:
                w_relax.append(bool(mf["relax_gamma"][pos0[hits[0]]])); h_relax.append(bool(mf["relax_H"][pos0[hits[0]]]))
    if do_amp:
        h_name = "Amplitude_l0_rgb"
        h_inputs = [float(PI_LD * LD(w_inputs[k]) * LD(h_inputs[k])) for k in range(len(h_inputs))]
    else:
        h_name = "Height_l0_rgb"
    height = Block(len(h_relax))
    width = Block(1 if dwa == 1 else 6)

    # ---- the bias spline: nodes and values from the hyper-prior rows (:254-285) ----
    hp, hpn = np.asarray(mf["hyper_priors"], dtype=np.float64), list(mf["hyper_priors_names"])
    nrows = hp.shape[0]
    Nfix = 0
    for i in range(nrows - 1):
        if hp[i + 1, 0] < hp[i, 0]:
            raise ValueError("the reference frequencies of the bias spline must increase (io_asymptotic.cpp:259-262)")
        if hpn[i] == "Fix":
            Nfix += 1
    if Nfix != len(hpn) - 1 and Nfix != 0:
        raise ValueError("either all or none of the bias values may be fixed (io_asymptotic.cpp:267-270)")
    fref = hp[:, 0].copy()
    ferr = np.zeros(nrows) if hp.shape[1] == 1 else hp[:, 1].copy()
    Nmm = NG + 2 * nrows + 1
    freq = Block(Nf_el[0] + Nmm + Nf_el[2] + Nf_el[3])
    tmp = [Hmin, Hmax, EMPTY, EMPTY]
    for k in range(len(h_inputs)):
        height.fill(h_name, "Jeffreys" if h_relax[k] else "Fix", h_inputs[k], tmp, k, 0)
    cpt = 0
    tmp = [EMPTY] * 4                      # (the reference's tmpXd still holds the height bounds when a fixed frequency is filled first: "Fix" ignores it)
    for k in range(len(f_inputs)):
        if k < Nf_el[0] or k >= Nf_el[0] + Nf_el[1]:
            if f_relax[k]:
                tmp = [f_min[k], f_max[k], 0.0025 * Dnu, 0.0025 * Dnu]
                freq.fill("Frequency_RGB_l", "GUG", f_inputs[k], tmp, cpt, 0)
            else:
                freq.fill("Frequency_RGB_l", "Fix", f_inputs[k], tmp, cpt, 0)
            cpt += 1
        elif k == Nf_el[0]:
            cpt += Nmm
    cpt = Nf_el[0] + NG + 1
    for k in range(nrows):
        freq.fill("fref_bias", "Fix", fref[k], [EMPTY] * 4, cpt, 0)
        cpt += 1
    if hp.shape[1] == 1:
        freq.fill("ferr_bias", "Uniform", ferr[0], [-Dnu / 2, Dnu / 20, EMPTY, EMPTY], cpt, 0); cpt += 1
        for k in range(1, nrows - 1):
            freq.fill("ferr_bias", "Uniform", ferr[k], [-Dnu / 20, Dnu / 20, EMPTY, EMPTY], cpt, 0); cpt += 1
        freq.fill("ferr_bias", "Uniform", ferr[nrows - 1], [-Dnu / 20, Dnu / 2, EMPTY, EMPTY], cpt, 0); cpt += 1
    else:
        for k in range(nrows):
            tmp = [EMPTY] * 4
            for c in range(hp.shape[1] - 2):
                tmp[c] = hp[k, 2 + c]
            freq.fill("ferr_bias", hpn[k], ferr[k], tmp, cpt, 0); cpt += 1
    if numax <= 0:
        numax = _getnumax(freq.inputs[:Nf_el[0]], height.inputs)
    elif err_numax <= 0:
        err_numax = 0.05 * numax
    snlm = Block(10)
    extra = np.array([1, 2., 0.2, 0, 3], dtype=np.float64)           # :419-432 (extra_priors[4] = 3 for both v4 models)

    trunc_c, model_type, bias_type = -1.0, -1, -1
    nsplit = 0
    for i, nm in enumerate(names):
        pr, row = cpri[i], mc[i]
        if nm in ("freq_smoothness", "Freq_smoothness"):
            if pr != "bool":
                _fatal("freq_smoothness", "bool")
            extra[0], extra[1] = row[0], row[1]
        if nm == "trunc_c":
            if pr != "Fix":
                _fatal("trunc_c", "Fix")
            trunc_c = row[0]
        if nm == "model_type":
            if pr != "Fix":
                _fatal("model_type", "Fix")
            model_type = int(row[0])
        if nm == "bias_type":
            if pr != "Fix":
                _fatal("bias_type", "Fix")
            bias_type = int(row[0]) if Nfix != len(hpn) - 1 else 0
        if nm in ("Frequency", "frequency"):
            if pr not in ("GUG", "Uniform"):
                _fatal(nm, "GUG or Uniform")
            pe = 0
            for k in range(len(f_inputs)):
                if k < Nf_el[0] or k >= Nf_el[0] + Nf_el[1]:
                    t4 = [f_min[k], f_max[k], row[3], row[4]] if pr == "GUG" else [f_min[k], f_max[k], EMPTY, EMPTY]
                    freq.fill("Frequency_l", pr if f_relax[k] else "Fix", f_inputs[k], t4, pe, 0)
                    pe += 1
                elif k == Nf_el[0]:
                    pe += Nmm
        if nm == "delta01":
            if pr == "Fix_Auto":
                freq.fill("delta01", "Uniform", 0.5 * Dnu / 100, [-1. * Dnu / 100, 1. * Dnu / 100, EMPTY, EMPTY], Nf_el[0], 0)
            else:
                freq.fill("delta01", pr, row[0], row, Nf_el[0], 1)
        if nm in _RGB_L1_POS:
            if pr == "Fix_Auto":
                _fatal(nm, "Fix_Auto")
            freq.fill(nm, pr, row[0], row, Nf_el[0] + _RGB_L1_POS[nm], 1)
        if nm in ("height", "Height", "amplitude", "Amplitude"):
            if pr == "Fix_Auto":
                _fatal(nm, "Fix_Auto")
            for k in range(len(h_inputs)):
                if h_relax[k]:
                    height.fill(h_name, pr, h_inputs[k], row, k, 0)
                else:
                    height.fill(h_name, "Fix", h_inputs[k], row, k, 1)
        if nm in ("width", "Width") and dwa == 1:
            if pr != "Fix_Auto":
                raise ValueError("the constant-width model takes Width Fix_Auto only (io_asymptotic.cpp:637-641)")
            mean = 0.0
            for k in range(len(w_inputs)):
                mean = mean + w_inputs[k] / len(w_inputs)
            width.fill("Width_l", "Jeffreys", mean, [resol, Dnu / 3., EMPTY, EMPTY], 0, 0)
        if nm in _RGB_SPLIT_POS:
            nsplit += 1
            if pr == "Fix_Auto":
                raise ValueError("Fix_Auto requested for %s: not allowed" % nm)
            p0 = _RGB_SPLIT_POS[nm]
            snlm.fill(_RGB_SPLIT_NAME[p0], pr, row[0], row, p0, 1)
        if nm in ("asphericity_eta", "Asphericity_eta"):
            snlm.names[8] = "eta0_switch"; snlm.pnames[8] = "Fix"; snlm.relax[8] = 0; snlm.inputs[8] = 0
        if nm == "eta0_switch":
            if pr != "Fix":
                raise ValueError("eta0_switch must be Fix 0 or 1 (io_asymptotic.cpp:973-981)")
            snlm.fill("eta0_switch", pr, row[0], row, 8, 1)
        if nm in ("asymetry", "Asymetry"):
            if pr == "Fix_Auto":
                _fatal("asymetry", "Fix_Auto")
            snlm.fill("Lorentzian_asymetry", pr, row[0], row, 9, 1)
        for l in (1, 2, 3):
            if nm in ("visibility_l%d" % l, "Visibility_l%d" % l):
                if pr == "Fix_Auto":
                    _fatal("visibility_l%d" % l, "Fix_Auto")
                if lmax >= l:
                    vis.fill("Visibility_l%d" % l, pr, row[0], row, l - 1, 1)
        if nm in ("inclination", "Inclination"):
            if pr == "Fix_Auto":
                _fatal("inclination", "Fix_Auto")
            inc.fill("Inclination", pr, 89.99999 if row[0] >= 90 else row[0], row, 0, 1)
    if nsplit != 8:
        raise ValueError("set rot_env, rot_core, a2_core, a2_env, a3_env, a4_env, a5_env, a6_env (8 parameters) (io_asymptotic.cpp:741-745)")
    noise = set_noise_params(np.asarray(mf["noise_s2"], dtype=np.float64), mf["noise_params"])
    if (model_type == -1) != (bias_type == -1):
        raise ValueError("model_type and bias_type must be given together (io_asymptotic.cpp:772-776)")
    if dwa == 2:
        width = set_width_app2016_params_v2(numax, err_numax)
    ncfg = 3 if (model_type == -1 and bias_type == -1) else 6
    plength = np.array([len(h_inputs), lmax, Nf_el[0], Nmm, Nf_el[2], Nf_el[3], len(snlm), len(width), len(noise), len(inc), ncfg], dtype=np.int64)
    allp = Block(int(plength.sum()))
    p0 = 0
    for blk in (height, vis, freq, snlm, width, noise, inc):
        n = len(blk)
        allp.names[p0:p0 + n] = blk.names
        allp.pnames[p0:p0 + n] = blk.pnames
        allp.inputs[p0:p0 + n] = blk.inputs
        allp.relax[p0:p0 + n] = blk.relax
        allp.priors[:, p0:p0 + n] = blk.priors
        p0 += n
    allp.fill("Truncation parameter", "Fix", trunc_c, mc[0], p0, 1)
    if allp.inputs[p0] <= 0:
        allp.inputs[p0] = 10000.
    allp.fill("Switch for fit of Amplitudes or Heights", "Fix", do_amp, mc[0], p0 + 1, 1)
    allp.fill("Maximum limit on random values generated by N(0,sigma_m)", "Fix", sigma_limit, [EMPTY] * 5, p0 + 2, 1)
    if model_type != -1:
        allp.fill("model type ", "Fix", model_type, [EMPTY] * 5, p0 + 3, 1)
    if bias_type != -1:
        allp.fill("bias type ", "Fix", bias_type, [EMPTY] * 5, p0 + 4, 1)
        allp.fill("Nferr ", "Fix", nrows, [EMPTY] * 5, p0 + 5, 1)
    if bias_type == 0 and Nfix != len(hpn) - 1:
        for k in [j for j, nm in enumerate(allp.names) if nm == "ferr_bias"]:
            allp.fill("ferr_bias", "Fix", 0, [EMPTY] * 5, k, 1)
    return {"model_fullname": fullname, "inputs": allp.inputs, "relax": allp.relax, "priors": allp.priors, "plength": plength,
            "extra_priors": extra, "inputs_names": allp.names, "priors_names": allp.pnames, "numax": numax, "err_numax": err_numax}


def _proj(a1, inc_deg, fn):
    """sqrt(a1) * cos / sin (inc * pi / 180) as io_ms_global.cpp:1204-1216 evaluates it: the angle and its cosine in long double"""
    return float(LD(math.sqrt(a1)) * fn(LD(inc_deg) * PI_LD / LD(180.)))


def _getnumax(fl, Hl):
    """io_ms_global.cpp:1559-1574: height-weighted mean frequency"""
    num = 0.0
    for i in range(len(fl)):
        num = num + fl[i] * Hl[i]
    return num / float(np.sum(Hl))


# model_fullname -> model id of include/tamcmc_gpu.h (the case labels of Model_def::call_model, model_def.cpp:220-388)
GPU_MODEL_IDS = {
    "model_MS_Global_a1etaa3_HarveyLike_Classic": 3, "model_MS_Global_a1etaa3_HarveyLike_Classic_v2": 12,
    "model_MS_Global_a1etaa3_HarveyLike_Classic_v3": 13, "model_MS_Global_a1l_etaa3_HarveyLike": 6, "model_MS_Global_a1n_etaa3_HarveyLike": 7,
    "model_MS_Global_a1nl_etaa3_HarveyLike": 8, "model_MS_Global_ajAlm_HarveyLike": 21, "model_MS_Global_aj_HarveyLike": 23,
}


# ================================================================================================================================
# The LOCAL-fit dialect: build_init_local (tamcmc/sources/io_local.cpp:329-1176) with set_noise_params_local (:1178-1238)
# ================================================================================================================================
LOCAL_MODELS = {"model_MS_local_basic": 11, "model_MS_local_Hnlm": 14}       # model_fullname -> id of tamcmc_gpu_create (models_ctrl.list)


def _fatal_local(var, kind):
    raise ValueError("%s: %s (fatalerror_msg_io_local, io_local.cpp:1240-1260)" %
                     (var, "Fix_Auto is not implemented for that parameter" if kind == "Fix_Auto" else "should always be defined as '%s'" % kind))


def harvey_like(noise_params, x):
    """tamcmc/sources/noise_models.cpp:15-39 on a zero spectrum: sum of Harvey-like profiles + white noise at the frequencies x"""
    x = np.asarray(x, dtype=np.float64)
    out = np.zeros_like(x)
    nh = (len(noise_params) - 1) // 3
    for k in range(nh):
        H, tc, p = noise_params[3 * k:3 * k + 3]
        if tc != 0:
            out = out + H * (1.0 / (np.array([math.pow(v, p) for v in (1e-3) * tc * x]) + 1.0))        # (libm's pow, like the reference)
    return out + noise_params[-1]


def set_noise_params_local(noise_params, freq_range):
    """io_local.cpp:1178-1238: a local fit approximates the noise by a constant -- the mean of the file's noise model at the two ends of
    the analysed range, Uniform between half its minimum and 1.5 times its maximum there."""
    nz = Block(1)
    nz.names[0], nz.pnames[0], nz.relax[0] = "White_Noise_N0", "Uniform", 1
    noise_params = np.asarray(noise_params, dtype=np.float64)
    n_skip = int(np.sum(np.abs(noise_params - (-2.0)) <= 1e-6))           # where_dbl(noise_params, -2, 1e-6)
    noise_p = np.full(len(noise_params) - n_skip, EMPTY)
    cpt = 0
    for v in noise_params:
        if v == -1:
            noise_p[cpt] = 0
            cpt += 1
        if v >= 0:
            noise_p[cpt] = v
            cpt += 1
    vals = harvey_like(noise_p, np.array([freq_range[0], freq_range[1]], dtype=np.float64))
    nz.inputs[0] = vals.sum() / len(vals)
    nz.priors[0, 0] = vals.min() * 0.5
    nz.priors[1, 0] = vals.max() * 1.5
    return nz


def build_init_local(mf, resol):
    """mf: the dictionary of formats.read_local_model (one slice of the file); resol: the spectrum's resolution.  Same return value as
    build_init_ms_global.  The models are model_MS_local_basic (id 11) and model_MS_local_Hnlm (id 14): every mode of the slice has its
    own height, width and frequency; no visibilities; the noise is one constant."""
    Hmin, Hmax = 1.0, 10000.0
    G = 6.667e-8
    Dnu_sun, R_sun, M_sun = 135.1, 6.96342e5, 1.98855e30
    rho_sun = float(LD(M_sun * 1e3) / ((LD(4) * PI_LD * LD(math.pow(R_sun * 1e5, 3))) / LD(3)))          # :337 (long double pi)
    Dnu = mf["Dnu"] if mf["Dnu"] is not None else -9999.0
    rho = math.pow(Dnu / Dnu_sun, 2.) * rho_sun
    names, cpri, mc = mf["common_names"], mf["common_names_priors"], mf["modes_common"]
    els = np.asarray(mf["els"])
    lmax = int(els.max())
    fr0, fr1 = mf["freq_range"]
    trunc_c = -1.0

    # ---- instructions that come before the set-up (:377-406) ----
    fullname, do_amp = " ", 0
    do_a11_eq_a12, do_avg_a1n = 1, 1
    for i, nm in enumerate(names):
        if nm == "model_fullname":
            fullname = cpri[i]
        if nm == "fit_squareAmplitude_instead_Height":
            if cpri[i] != "bool":
                _fatal_local(nm, "bool")
            do_amp = int(bool(mc[i, 0]))
    if fullname == " ":
        raise ValueError("Model name empty: the .model file needs the model_fullname variable (io_local.cpp:402-405)")
    hnlm = fullname == "model_MS_local_Hnlm"
    inc = Block(1)

    # ---- frequencies / widths / heights of the eigen table, degree by degree, matched with the relax list (:414-503) ----
    eig = np.asarray(mf["eigen_params"], dtype=np.float64)
    f_in, h_in, w_in, fmin_in, fmax_in = [[] for _ in range(4)], [[] for _ in range(4)], [[] for _ in range(4)], [[] for _ in range(4)], [[] for _ in range(4)]
    f_rl, h_rl, w_rl = [[] for _ in range(4)], [[] for _ in range(4)], [[] for _ in range(4)]
    for el in range(lmax + 1):
        pos0 = [k for k in range(len(els)) if els[k] == el]
        if not pos0:
            continue
        f_el = [mf["freqs_ref"][k] for k in pos0]
        pos_el = [k for k in range(eig.shape[0]) if int(eig[k, 0]) == el]
        for k in pos_el:
            if el <= 3:
                f_in[el].append(eig[k, 1]); fmin_in[el].append(eig[k, 2]); fmax_in[el].append(eig[k, 3]); w_in[el].append(eig[k, 4]); h_in[el].append(eig[k, 5])
            hits = [j for j in range(len(f_el)) if eig[k, 1] - 1e-2 <= f_el[j] <= eig[k, 1] + 1e-2]      # where_dbl, string_handler.cpp:88-110
            if len(hits) != 1:
                raise ValueError("the frequency %r is not unique in / absent from the relax list (io_local.cpp:484-497)" % eig[k, 1])
            if el <= 3:
                f_rl[el].append(bool(mf["relax_freq"][pos0[hits[0]]])); w_rl[el].append(bool(mf["relax_gamma"][pos0[hits[0]]]))
                h_rl[el].append(bool(mf["relax_H"][pos0[hits[0]]]))

    # ---- only the modes strictly inside the analysed range (filter_range, string_handler.cpp:515-557; :510-560) ----
    for el in range(4):
        if not f_in[el]:
            continue
        keep = [k for k, f in enumerate(f_in[el]) if fr0 < f < fr1]
        for lst in (h_in, w_in, fmin_in, fmax_in, f_rl, h_rl, w_rl, f_in):       # (f_in last: it is the filter's key)
            lst[el] = [lst[el][k] for k in keep]
    Nf_el = [len(f_in[el]) for el in range(4)]
    if sum(len(h) for h in h_in) == 0:
        raise ValueError("No parameters found in the specified frequency range (io_local.cpp:562-569)")

    # ---- heights / amplitudes, widths, frequencies: the defaults (:574-668) ----
    if do_amp:
        h_name = "Amplitude_l"
        for el in range(4):
            h_in[el] = [float(PI_LD * LD(w_in[el][k]) * LD(h_in[el][k])) for k in range(len(h_in[el]))]
    else:
        h_name = "Height_l"
    tmp = [Hmin, Hmax, EMPTY, EMPTY]

    def fill_vect(blk, vals, relax, name, prior, prior_vals, pos, i0_if, i0_else):
        """IO_models::fill_param_vect / fill_param_vect2 (io_models.cpp:79-118): prior_vals is one row for all, or one row per value"""
        per_value = len(prior_vals) > 0 and hasattr(prior_vals[0], "__len__")
        for k in range(len(vals)):
            row = prior_vals[k] if per_value else prior_vals
            if relax[k]:
                blk.fill(name, prior, vals[k], row, k + pos, i0_if)
            else:
                blk.fill(name, "Fix", vals[k], row, k + pos, i0_else)

    offs = [0, len(h_in[0]), len(h_in[0]) + len(h_in[1]), len(h_in[0]) + len(h_in[1]) + len(h_in[2])]
    if hnlm:
        height = Block(Nf_el[0] + Nf_el[1] * 2 + Nf_el[2] * 3 + Nf_el[3] * 4)
        p0 = 0
        fill_vect(height, h_in[0], h_rl[0], h_name, "Jeffreys", tmp, p0, 0, 0)
    else:
        height = Block(sum(Nf_el))
        for el in range(4):
            p0 = offs[el]
            fill_vect(height, h_in[el], h_rl[el], h_name, "Jeffreys", tmp, p0, 0, 0)
    width, freq = Block(sum(Nf_el)), Block(sum(Nf_el))
    tmp = [resol, Dnu / 3., EMPTY, EMPTY] if Dnu > 0 else [resol, 20., EMPTY, EMPTY]
    for el in range(4):
        p0 = offs[el]
        fill_vect(width, w_in[el], w_rl[el], "Width_l", "Jeffreys", tmp, p0, 0, 0)
    for el in range(4):
        p0 = offs[el]
        for k in range(len(f_in[el])):
            if f_rl[el][k]:
                d = 0.01 * abs(fmax_in[el][k] - fmin_in[el][k])
                tmp = [fmin_in[el][k], fmax_in[el][k], d, d]
                freq.fill("Frequency_l", "GUG", f_in[el][k], tmp, k + p0, 0)
            else:
                freq.fill("Frequency_l", "Fix", f_in[el][k], tmp, k + p0, 0)
    snlm = Block(6)                                                               # do_a11_eq_a12 == do_avg_a1n == 1 for both models (:671-673)
    extra = np.array([0, 0, 0.2, 0], dtype=np.float64)                            # :693-698

    # ---- the common parameters (:700-968) ----
    pos_prior_height = -1
    bool_a1sini = bool_a1cosi = False
    for i, nm in enumerate(names):
        pr, row = cpri[i], mc[i]
        if nm == "trunc_c":
            if pr != "Fix":
                _fatal_local("trunc_c", "Fix")
            trunc_c = row[0]
        if nm in ("height", "Height", "amplitude", "Amplitude"):
            amp = nm in ("amplitude", "Amplitude")
            if pr == "Fix_Auto":
                def auto_rows(hs):
                    if amp:
                        return [[float(PI_LD * LD(Dnu) / LD(3.) * LD(h) / LD(row[0])), float(PI_LD * LD(Dnu) / LD(3.) * LD(h) * LD(row[1])), EMPTY, EMPTY] for h in hs]
                    return [[h / row[0], h * row[1], EMPTY, EMPTY] for h in hs]
                if hnlm:
                    pos_prior_height = i
                    fill_vect(height, h_in[0], h_rl[0], h_name, "Jeffreys", auto_rows(h_in[0]), p0, 0, 0)
                else:
                    for el in range(4):
                        p0 = offs[el]
                        fill_vect(height, h_in[el], h_rl[el], h_name, "Jeffreys", auto_rows(h_in[el]), p0, 0, 0)
            else:
                if hnlm:
                    pos_prior_height = i
                    fill_vect(height, h_in[0], h_rl[0], h_name, pr, row, p0, 0, 0)
                else:
                    for el in range(4):
                        p0 = offs[el]
                        fill_vect(height, h_in[el], h_rl[el], h_name, pr, row, p0, 1, 1)
        if nm in ("width", "Width"):
            if pr == "Fix_Auto":
                wname = "Jeffreys"
                tmp = [resol, Dnu / 3, EMPTY, EMPTY] if Dnu > 0 else [resol, 20, EMPTY, EMPTY]
            else:
                wname = pr
                tmp = list(row)
            for el in range(4):
                p0 = offs[el]
                fill_vect(width, w_in[el], w_rl[el], "Width_l", wname, tmp, p0, 0, 1)
        if nm in ("splitting_a1", "Splitting_a1"):
            if pr == "Fix_Auto":
                _fatal_local("splitting_a1", "Fix_Auto")
            p0 = 0
            snlm.fill("Splitting_a1", pr, row[0], row, p0, 1)
        if nm in ("asphericity_eta", "Asphericity_eta"):
            snlm.names[1] = "Asphericity_eta0"
            if pr == "Fix_Auto":
                snlm.pnames[1] = "Fix"
                snlm.relax[1] = 0
                if row[0] == 1:
                    snlm.inputs[1] = 3. / (4. * math.pi * rho * G) if Dnu > 0 else 0.0
                else:
                    snlm.inputs[1] = 0
            else:
                p0 = 1
                snlm.fill("Asphericity_eta", pr, row[0], row, p0, 1)
        if nm in ("splitting_a3", "Splitting_a3"):
            if pr == "Fix_Auto":
                _fatal_local("splitting_a3", "Fix_Auto")
            p0 = 2
            snlm.fill("Splitting_a3", pr, row[0], row, p0, 1)
        if nm in ("asymetry", "Asymetry"):
            if pr == "Fix_Auto":
                _fatal_local("asymetry", "Fix_Auto")
            p0 = 5
            snlm.fill("Lorentzian_asymetry", pr, row[0], row, p0, 1)
        if nm in ("inclination", "Inclination"):
            if pr == "Fix_Auto":
                _fatal_local("inclination", "Fix_Auto")
            p0 = 0
            inc.fill("Inclination", pr, 89.99999 if row[0] >= 90 else row[0], row, p0, 1)
        if nm == "sqrt(splitting_a1).cosi":
            if pr == "Fix_Auto":
                _fatal_local(nm, "Fix_Auto")
            p0 = 3
            snlm.fill(nm, pr, row[0], row, p0, 1)
            bool_a1cosi = True
        if nm == "sqrt(splitting_a1).sini":
            if pr == "Fix_Auto":
                _fatal_local(nm, "Fix_Auto")
            p0 = 4
            snlm.fill(nm, pr, row[0], row, p0, 1)
            bool_a1sini = True

    # ---- splitting_a1 + inclination -> sqrt(a1).cosi, sqrt(a1).sini (:971-1063) ----
    if bool_a1cosi != bool_a1sini:
        raise ValueError("both 'sqrt(splitting_a1).sini' and 'sqrt(splitting_a1).cosi' must appear (io_local.cpp:971-977)")
    if not bool_a1cosi:
        if fullname == "model_MS_local_basic":
            ang = LD(inc.inputs[0]) * PI_LD / LD(180.)
            c, s_ = float(LD(math.sqrt(snlm.inputs[0])) * np.cos(ang)), float(LD(math.sqrt(snlm.inputs[0])) * np.sin(ang))
            if inc.pnames[0] == "Fix" and snlm.pnames[0] == "Fix":
                snlm.fill("sqrt(splitting_a1).cosi", "Fix", c, snlm.priors[:, 0].copy(), 3, 0)
                snlm.fill("sqrt(splitting_a1).sini", "Fix", s_, snlm.priors[:, 0].copy(), 4, 0)
            else:
                snlm.priors[1, 0] = math.sqrt(snlm.priors[1, 0])
                snlm.fill("sqrt(splitting_a1).cosi", snlm.pnames[0], c, snlm.priors[:, 0].copy(), 3, 0)
                snlm.fill("sqrt(splitting_a1).sini", snlm.pnames[0], s_, snlm.priors[:, 0].copy(), 4, 0)
            if snlm.inputs[3] < 1e-2:
                snlm.inputs[3] = 1e-2
            if snlm.inputs[4] < 1e-2:
                snlm.inputs[4] = 1e-2
            inc.fill("Empty", "Fix", 0, inc.priors[:, 0].copy(), 0, 1)
            snlm.fill("Empty", "Fix", 0, snlm.priors[:, 0].copy(), 0, 1)
        if hnlm:
            inc0 = inc.inputs[0]
            inc.fill("Empty", "Fix", 0, inc.priors[:, 0].copy(), 0, 1)
            ind = len(h_in[0])
            if pos_prior_height <= 0:
                tmp, tname = [Hmin, Hmax, EMPTY, EMPTY], "Jeffreys"
            else:
                tmp, tname = list(mc[pos_prior_height]), cpri[pos_prior_height]
            for el in range(1, lmax + 1):
                r = amplitude_ratio(el, inc0)
                for en in range(Nf_el[el]):
                    for em in range(el + 1):
                        height.fill("H(%d,%d,%d)" % (en, el, em), tname, h_in[el][en] * r[el + em], tmp, ind, 0)
                        ind += 1
            extra[3] = 2
    else:
        inc.fill("Empty", "Fix", 0, inc.priors[:, 0].copy(), 0, 1)
        snlm.fill("Empty", "Fix", 0, snlm.priors[:, 0].copy(), 0, 1)

    noise = set_noise_params_local(mf["noise_params"], (fr0, fr1))

    # ---- everything in one vector (:1076-1140) ----
    nh = len(h_in[0]) + 2 * len(h_in[1]) + 3 * len(h_in[2]) + 4 * len(h_in[3]) if hnlm else sum(len(h) for h in h_in)
    plength = np.array([nh, 0, Nf_el[0], Nf_el[1], Nf_el[2], Nf_el[3], len(snlm), sum(len(w) for w in w_in), len(noise), len(inc), 2], dtype=np.int64)
    allp = Block(int(plength.sum()))
    p0 = 0
    for blk in (height, freq, snlm, width, noise, inc):
        n = len(blk)
        allp.names[p0:p0 + n] = blk.names
        allp.pnames[p0:p0 + n] = blk.pnames
        allp.inputs[p0:p0 + n] = blk.inputs
        allp.relax[p0:p0 + n] = blk.relax
        allp.priors[:, p0:p0 + n] = blk.priors
        p0 += n
    row0 = mc[0]
    allp.fill("Truncation parameter", "Fix", trunc_c, row0, p0, 1)
    if allp.inputs[p0] <= 0:
        allp.inputs[p0] = 10000.
    allp.fill("Switch for fit of Amplitudes or Heights", "Fix", do_amp, row0, p0 + 1, 1)
    return {"model_fullname": fullname, "inputs": allp.inputs, "relax": allp.relax, "priors": allp.priors, "plength": plength,
            "extra_priors": extra, "inputs_names": allp.names, "priors_names": allp.pnames, "numax": mf["numax"], "err_numax": mf.get("err_numax", EMPTY)}
