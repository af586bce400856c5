"""Readers for the reference's simplest on-disk inputs and reader/writer of its binary chain outputs (first slice): the ASCII `.data` spectrum and
the "simple matrix" `.model` file of the Gaussian-envelope fits (model_Harvey_Gaussian / model_Kallinger2014_Gaussian).

  .data  : '#' comments, '!' column labels, '*' units, then whitespace-separated columns (frequency, power[, ...])
           -- Config::read_data_ascii_Ncols (tamcmc/sources/config.cpp)
  .model : '#' comments, '* fmin fmax' the fitted range, '! names', the initial values, '! relax' + one 0/1 flag per
           parameter, '! prior names', then up to four rows of prior parameters (-9999 = unused slot)
           -- Config::read_inputs_prior_Simple_Matrix (config.cpp:560-660)

  .cfg   : '!Group:' headers, `key=value; free text` lines (Config::read_cfg_file / format_line, config.cpp:1062-1110, 1223-1500)
  outputs: <prefix>.hdr + <prefix>_chain-<k>.bin, the binary chain files of Outputs::write_bin_params (outputs.cpp:1231-1334)

Host-side I/O only: nothing here is on the GPU path."""
import numpy as np

# Config/default/primepriors_ctrl.list
PRIOR_KINDS = {"None": 0, "Fix": 0, "Uniform": 1, "Gaussian": 2, "multivar_Gaussian": 3, "Jeffreys": 4, "UG": 5, "GU": 6, "GUG": 7,
               "Uniform_abs": 8, "Uniform_cos": 9, "Jeffreys_abs": 10, "Tabulated": 11, "Tabulated_2d": 12, "Auto": 13}


def read_data(path, x_col=0, y_col=1, xrange=None):
    """-> (x, y) of an ASCII spectrum, optionally cut to xrange = (fmin, fmax) like the reference does with the '*' line of
    the .model file (Data.xrange)."""
    rows = []
    with open(path) as f:
        for line in f:
            s = line.strip()
            if not s or s[0] in "#!*":
                continue
            rows.append([float(v) for v in s.split()])
    a = np.asarray(rows, dtype=np.float64)
    x, y = np.ascontiguousarray(a[:, x_col]), np.ascontiguousarray(a[:, y_col])
    if xrange is not None:
        keep = (x >= xrange[0]) & (x <= xrange[1])
        x, y = np.ascontiguousarray(x[keep]), np.ascontiguousarray(y[keep])
    return x, y


def read_simple_matrix_model(path):
    """-> dict(xrange, names, inputs, relax, prior_names, prior_kinds, priors[4, Nparams])"""
    with open(path) as f:
        lines = [l.strip() for l in f if l.strip()]
    i = 0
    while i < len(lines) and lines[i].startswith("#"):
        i += 1
    if i >= len(lines) or not lines[i].startswith("*"):
        raise ValueError("%s: no '*' line with the frequency range" % path)
    xrange = [float(v) for v in lines[i][1:].split()]
    i += 1
    if not lines[i].startswith("!"):
        raise ValueError("%s: no '!' line with the parameter names" % path)
    names = lines[i][1:].split()
    inputs = np.array([float(v) for v in lines[i + 1].split()])
    if not lines[i + 2].startswith("!"):
        raise ValueError("%s: no '! relax' line" % path)
    relax = np.array([int(float(v)) for v in lines[i + 3].split()], dtype=np.int32)
    if not lines[i + 4].startswith("!"):
        raise ValueError("%s: no '!' line with the prior names" % path)
    prior_names = lines[i + 4][1:].split()
    n = len(prior_names)
    priors = np.full((4, n), -9999.0)
    for r, l in enumerate(lines[i + 5:i + 9]):
        v = [float(t) for t in l.split()]
        priors[r, :len(v)] = v
    if not (len(names) == len(inputs) == len(relax) == n):
        raise ValueError("%s: %d names, %d values, %d relax flags, %d priors" % (path, len(names), len(inputs), len(relax), n))
    kinds = np.array([PRIOR_KINDS.get(p, -1) for p in prior_names], dtype=np.int32)
    return dict(xrange=xrange, names=names, inputs=inputs, relax=relax, prior_names=prior_names, prior_kinds=kinds, priors=priors)


# ------------------------------------------------------------------------------------------------
# Binary chain outputs of the reference (Outputs::write_bin_params, tamcmc/sources/outputs.cpp:1231-1334):
#   <prefix>.hdr            ASCII metadata, '#' comments and '! key= values' lines
#   <prefix>_chain-<k>.bin  raw little-endian float64, one row of Nvars values per kept sample, one file per chain
# so that a run of the GPU driver leaves files the reference's own post-processing tools (bin2txt, getstats) read.
# ------------------------------------------------------------------------------------------------
def _eigen_row(values, integers):
    """A row vector the way Eigen's operator<< prints `v.transpose()`: default stream precision (6 significant digits),
    every coefficient right-aligned to the widest one, single-space separators."""
    txt = [("%d" % int(v)) if integers else ("%g" % float(v)) for v in values]
    w = max((len(t) for t in txt), default=0)
    return " ".join(t.rjust(w) for t in txt)


def params_header_text(Nsamples, Nchains, Nsamples_done, relax, plength, cons_names, cons_values, var_names):
    """The text of <prefix>.hdr (outputs.cpp:1259-1303).  `relax` has one flag per parameter (variables and constants),
    `cons_names == ['None']` stands for "no constant" (the reference then writes -1 as the value)."""
    Nvars, Ncons = len(var_names), len(cons_names)
    out = ["# This is the header file of the BINARY output file for the model parameters ",
           "# This file contains values for vars[0:Nchains-1][ 0:Nvars-1]. Each matrix is in a different file, indexed by the chain number",
           "! Nsamples= %d" % Nsamples, "! Nchains= %d" % Nchains, "! Nsamples_done=%d" % Nsamples_done, "! Nvars= %d" % Nvars,
           "! Ncons= %d" % Ncons, "! relax= " + _eigen_row(relax, True), "! plength= " + _eigen_row(plength, True),
           "! constant_names= " + "".join("%s   " % n for n in cons_names),
           "! constant_values= " + ("-1" if list(cons_names)[:1] == ["None"] else _eigen_row(cons_values, False)),
           "! variable_names=" + "".join("%s   " % n for n in var_names)]
    return "\n".join(out) + "\n"


def write_params_outputs(prefix, samples, relax, plength, cons_names, cons_values, var_names, Nsamples=None, append=False, file_ext="bin"):
    """samples[Nkept, Nchains, Nvars] -> <prefix>.hdr and <prefix>_chain-<k>.<file_ext>.  append=True adds rows to existing
    chain files and leaves the header alone, like the reference's buffered writes after the first one."""
    samples = np.ascontiguousarray(samples, dtype=np.float64)
    if samples.ndim != 3 or samples.shape[2] != len(var_names):
        raise ValueError("samples must be [Nkept, Nchains, Nvars]")
    Nkept, Nchains, _ = samples.shape
    if not append:
        with open(prefix + ".hdr", "w") as f:
            f.write(params_header_text(Nkept if Nsamples is None else Nsamples, Nchains, Nkept, relax, plength, cons_names, cons_values, var_names))
    for k in range(Nchains):
        with open("%s_chain-%d.%s" % (prefix, k, file_ext), "ab" if append else "wb") as f:
            f.write(np.ascontiguousarray(samples[:, k, :]).astype("<f8").tobytes())


def parse_params_header(text):
    """-> dict(Nsamples, Nchains, Nsamples_done, Nvars, Ncons, relax, plength, constant_names, constant_values, variable_names)"""
    h = {}
    for line in text.splitlines():
        s = line.strip()
        if not s.startswith("!"):
            continue
        key, _, val = s[1:].partition("=")
        key, toks = key.strip(), val.split()
        if key in ("Nsamples", "Nchains", "Nsamples_done", "Nvars", "Ncons"):
            h[key] = int(toks[0])
        elif key in ("relax", "plength"):
            h[key] = np.array([int(float(t)) for t in toks], dtype=np.int64)
        elif key == "constant_values":
            h[key] = np.array([float(t) for t in toks], dtype=np.float64)
        else:
            h[key] = toks
    return h


def read_params_outputs(prefix, chains=None, file_ext="bin"):
    """-> (header dict, samples[Nrows, Nchains_read, Nvars]); Nrows is what the chain files hold (the reference appends whole
    buffers, so it can be ahead of or behind the header's Nsamples_done)."""
    with open(prefix + ".hdr") as f:
        h = parse_params_header(f.read())
    ks = range(h["Nchains"]) if chains is None else chains
    cols = []
    for k in ks:
        a = np.fromfile("%s_chain-%d.%s" % (prefix, k, file_ext), dtype="<f8")
        if a.size % h["Nvars"]:
            raise ValueError("chain file %d does not hold whole rows of Nvars=%d values" % (k, h["Nvars"]))
        cols.append(a.reshape(-1, h["Nvars"]))
    n = min(c.shape[0] for c in cols)
    return h, np.stack([c[:n] for c in cols], axis=1)


# ------------------------------------------------------------------------------------------------
# Acceptance-rate log (Outputs::write_txt_acceptance, outputs.cpp:747-789) and restore files
# (Outputs::write_buffer_restore, outputs.cpp:863-1027; read back by Config::restore at start-up when do_restore=1)
# ------------------------------------------------------------------------------------------------
ACCEPTANCE_HEADER = ("# This is an output file for the acceptance rate. \n"
                     "# This file contains values for the acceptance_rate[0:Nchains-1] in function of the average sample position\n"
                     "# Averaging is done over Nbuffer \n")


def acceptance_text(xaxis, rates, with_header=True):
    """Rows `xaxis rate_0 .. rate_{Nchains-1}` the way the reference appends them (one Eigen-printed row per buffer)."""
    rates = np.atleast_2d(np.asarray(rates, dtype=np.float64))
    out = (ACCEPTANCE_HEADER + "! Nchains= %d\n" % rates.shape[1]) if with_header else ""
    for x, r in zip(np.atleast_1d(xaxis), rates):
        out += "%g %s\n" % (x, _eigen_row(r, False))
    return out


def read_acceptance(path):
    """-> (xaxis[Nrows], rates[Nrows, Nchains])"""
    rows = [[float(t) for t in l.split()] for l in open(path) if l.strip() and l.lstrip()[0] not in "#!"]
    a = np.asarray(rows, dtype=np.float64)
    return a[:, 0].copy(), a[:, 1:].copy()


def parse_restore_text(text):
    """One restore file -> dict.  Scalars (`Nchains`, `Nvars`, `iteration`), `variable_names`, and every `! key=` block as an
    array: values on the key's own line give a vector (sigmas), following rows a matrix [Nchains, Nvars] (vars, mus, ...),
    `*k` sub-blocks a stack [Nchains, Nvars, Nvars] (covarmats).  Values carry the 6 significant digits the reference prints."""
    out, key, rows, blocks = {}, None, [], None

    def close():
        nonlocal key, rows, blocks
        if key is not None:
            if blocks is not None:
                if rows:
                    blocks.append(rows)
                out[key] = np.asarray(blocks, dtype=np.float64)
            elif rows:
                out[key] = np.asarray(rows, dtype=np.float64)
        key, rows, blocks = None, [], None

    for line in text.splitlines():
        s = line.strip()
        if not s or s[0] == "#":
            continue
        if s[0] == "!":
            close()
            k, _, v = s[1:].partition("=")
            k, toks = k.strip(), v.split()
            if k in ("Nchains", "Nvars", "iteration"):
                out[k] = int(toks[0])
            elif k == "variable_names":
                out[k] = toks
            elif toks:
                out[k] = np.asarray([float(t) for t in toks], dtype=np.float64)
            else:
                key = k
        elif s[0] == "*":
            if blocks is None:
                blocks = []
            elif rows:
                blocks.append(rows)
            rows = []
        else:
            rows.append([float(t) for t in s.split()])
    close()
    return out


def read_restore(directory, star_id, phase="A"):
    """The three files `<star_id>_restore_<phase>_{1,2,3}.dat` of a run -> one dict: vars, vars_mean (file 1), sigmas, mus and
    their means (file 2), covarmats, covarmats_mean (file 3) -- what a restarted run needs (position, proposal scale, proposal
    mean and covariance of every chain)."""
    import os
    merged = {}
    for n in (1, 2, 3):
        with open(os.path.join(directory, "%s_restore_%s_%d.dat" % (star_id, phase, n))) as f:
            d = parse_restore_text(f.read())
        for k in ("Nchains", "Nvars", "iteration"):
            if k in merged and k in d and merged[k] != d[k]:
                raise ValueError("restore files disagree on %s" % k)
        merged.update(d)
    return merged


_RESTORE_WHAT = {
    1: ("# Contains the last values for the variables vars[0:Nchain-1]. vars_mean denotes averaged values of Nbuffer \n", "do_restore_[X]=1", "do_restore_proposal=1"),
    2: ("# Contains the last values of (a) sigmas[0:Nchains-1] and (b) mus[0:Nchains-1, 0:Nvars-1].  sigmas_mean and mus_mean denotes averaged values of Nbuffer\n", "do_restore=1", "do_restore=1"),
    3: ("# Contains the last value of the covariance matrix covarmats[0:Nchains-1, 0:Nvars-1, 0:Nvars-1]. covarmats_mean denotes the averaged values over Nbuffer\n", "do_restore=1", "do_restore=1"),
}


def restore_texts(state):
    """state (the dict read_restore returns) -> {1: text, 2: text, 3: text} in the layout of Outputs::write_buffer_restore
    (outputs.cpp:863-1027): every matrix row printed on its own (`.row(i)`: aligned per row), 6 significant digits."""
    names = "! variable_names=" + "".join("%s   " % n for n in state["variable_names"]) + "\n"

    def head(n):
        what, a, b = _RESTORE_WHAT[n]
        return ("# This is an output file containing what is required to restore a run to its last saved position \n"
                "# File number: %d \n" % n + what + "# Use this if you wish to: \n"
                "#       (1) complete a finished job that requires more samples ==> set erase_old_file=0 and %s \n" % a +
                "#       (2) restart a finished job by ignoring old samples (e.g. ignoring a Burn-in) ==> set erase_old_file=1 and %s \n" % b +
                "#       (3) terminate an unfinished job which failed to finished (e.g. due to computer unexpected shutdown) ==> set erase_old_file=0 and %s \n" % a +
                "! Nchains= %d\n! Nvars= %d\n! iteration=%d\n" % (state["Nchains"], state["Nvars"], state["iteration"]) + names)

    def rows(M):
        return "".join(_eigen_row(r, False) + "\n" for r in np.atleast_2d(M))

    t1 = head(1) + "! vars= \n" + rows(state["vars"]) + "! vars_mean= \n" + rows(state["vars_mean"])
    t2 = (head(2) + "! sigmas= " + _eigen_row(state["sigmas"], False) + "\n! mus= \n" + rows(state["mus"]) +
          "! sigmas_mean= " + _eigen_row(state["sigmas_mean"], False) + "\n! mus_mean= \n" + rows(state["mus_mean"]))
    t3 = head(3)
    for key in ("covarmats", "covarmats_mean"):
        t3 += "! %s= \n" % key
        for k, C in enumerate(state[key]):
            t3 += "*%d\n" % k + rows(C)
    return {1: t1, 2: t2, 3: t3}


def write_restore(directory, star_id, state, phase="A"):
    import os
    for n, text in restore_texts(state).items():
        with open(os.path.join(directory, "%s_restore_%s_%d.dat" % (star_id, phase, n)), "w") as f:
            f.write(text)


# ------------------------------------------------------------------------------------------------
# The .cfg control file (Config::read_cfg_file / format_line, tamcmc/sources/config.cpp:1062-1110, 1223-1500):
#   '!Group:' opens a group, '#' starts a comment line, every other line is `key=value; free text` -- the value ends at the
#   FIRST ';' (a line without one is an error in the reference), numbers are read with strtod semantics (leading number,
#   trailing text ignored: `lambda_temp=3.50 #1.70;` is 3.5), lists are comma separated.
# ------------------------------------------------------------------------------------------------
import re as _re

_NUM = _re.compile(r"\s*([-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?))")


def cfg_number(raw):
    m = _NUM.match(raw)
    if not m:
        raise ValueError("no number at the start of %r" % raw)
    return float(m.group(1))


def cfg_list(raw):
    return [cfg_number(t) for t in raw.split(",") if t.strip()]


def read_cfg(path):
    """-> {group: {key: raw value string}} with the reference's line rules; later duplicates of a key win, as they do when the
    reference assigns keyword after keyword."""
    groups, cur = {}, None
    with open(path) as f:
        for n, line in enumerate(f, 1):
            s = line.strip()
            if not s or s[0] == "#":
                continue
            if s[0] == "!":
                cur = groups.setdefault(s[1:].split(":")[0].strip(), {})
                continue
            if s == "/END":
                break
            if ";" not in s:
                raise ValueError("%s:%d: no ';' terminates the value (config.cpp:1090-1096)" % (path, n))
            body = s[: s.index(";")].strip()
            if "=" not in body or cur is None:
                continue
            key, _, val = body.partition("=")
            cur[key.strip()] = val.strip()
    return groups


def mala_config(groups):
    """The !MALA group with the field names of tamcmc::DriverConfig (host/mcmc_driver.hpp); `epsilon2` is the reference's
    MALA.epsi2 (config.cpp:1276-1279)."""
    g = groups["MALA"]
    out = {"Nchains": int(cfg_number(g["Nchains"])), "lambda_temp": cfg_number(g["lambda_temp"]), "c0": cfg_number(g["c0"]),
           "epsilon1": cfg_number(g["epsilon1"]), "epsi2": cfg_number(g["epsilon2"]), "A1": cfg_number(g["A1"]),
           "target_acceptance": cfg_number(g["target_acceptance"]), "dN_mixing": int(cfg_number(g["dN_mixing"])),
           "Nt_learn": [int(v) for v in cfg_list(g["Nt_learn"])], "periods_learn": [int(v) for v in cfg_list(g["periods_learn"])]}
    if len(out["periods_learn"]) != len(out["Nt_learn"]) - 1:
        raise ValueError("periods_learn must have one entry fewer than Nt_learn (config_default.cfg:18)")
    return out


# ------------------------------------------------------------------------------------------------
# The MS_Global `.model` file, PARSE STAGE ONLY (read_MCMC_file_MS_Global, tamcmc/sources/io_ms_global.cpp:27-360): the file
# into the fields of the reference's MCMC_files structure.  The red-giant (asymptotic) dialect goes through the same reader
# (read_MCMC_file_asymptotic, io_asymptotic.cpp:27-29); there the hyper-prior rows `value  prior  params...` hold the l=1
# frequency priors of model_RGB_asympt_aj_AppWidth_HarveyLike_v4.  Turning those into the parameter vector / plength / priors
# (build_init_MS_Global, io_ms_global.cpp:362-1400) stays in the reference.
# The reference walks the file by counting lines that start with '#':
#   up to the 3rd '#'   header ('#KIC=' id, '! Dnu', '!! C_l', '!n numax [err]', '* fmin fmax') and the mode table
#                       `p|g|co  l  frequency  [relax_f relax_H relax_W]` (missing flags default to 1)
#   to the next '#'     hyper priors (the line right after the 3rd '#' is skipped unread -- in the shipped files that is the
#                       '# Extra parameters (obselete)' label, so the "extra parameters" column lands in hyper_priors)
#   then, '#' to '#':   eigen-solution table [n, 6], noise parameters (flattened, right-aligned in 10 slots, -1 fill),
#                       noise_s2 [<=10, 3] (right-aligned in 10 rows, -1 fill), and to the end of the file the common
#                       parameters `name  prior_or_switch  values...` (up to 5 values, -9999 fill)
# ------------------------------------------------------------------------------------------------
def read_ms_global_model(path, slice_ind=None):
    """slice_ind = None: the MS_Global / red-giant dialect (one '*' frequency range).  slice_ind = k: the LOCAL-fit dialect
    (read_MCMC_file_local, io_local.cpp:25-327): the same walk, but the file lists one '*' range per slice and the k-th one is the
    range that is analysed (io_local.cpp:76-88)."""
    with open(path) as f:
        lines = [l.strip() for l in f.read().splitlines()]
    range_counter = 0
    out = {"ID": None, "Dnu": None, "C_l": None, "numax": -9999.0, "err_numax": -9999.0, "freq_range": None,
           "param_type": [], "els": [], "freqs_ref": [], "relax_freq": [], "relax_H": [], "relax_gamma": []}
    i, nhash = 0, 0
    while nhash < 3 and i < len(lines):
        s = lines[i]
        i += 1
        if not s:
            continue
        if s[0] == "#":
            nhash += 1
            if s[1:2] == "K":
                out["ID"] = s.replace("=", " ").split()[1]
        elif s[0] == "!":
            w = s.split()
            if s[1:2] == "!":
                out["C_l"] = float(w[1])
            elif s[1:2] == "n":
                out["numax"] = float(w[1])
                if len(w) == 3:
                    out["err_numax"] = float(w[2])
            else:
                out["Dnu"] = float(w[1])
        elif s[0] == "*":
            w = s.split()
            if slice_ind is None:
                if out["freq_range"] is not None:
                    raise ValueError("two frequency ranges (io_ms_global.cpp: only one '*' line is allowed)")
                out["freq_range"] = (float(w[1]), float(w[2]))
            else:
                if range_counter == slice_ind:
                    out["freq_range"] = (float(w[1]), float(w[2]))
                range_counter += 1
        else:
            w = s.split()
            if w[0] not in ("p", "g", "co"):
                raise ValueError("mode line must start with p, g or co: %r" % s)
            out["param_type"].append(w[0]); out["els"].append(int(w[1])); out["freqs_ref"].append(float(w[2]))
            flags = [bool(int(float(t))) for t in w[3:6]] + [True] * (3 - len(w[3:6]))
            out["relax_freq"].append(flags[0]); out["relax_H"].append(flags[1]); out["relax_gamma"].append(flags[2])
    i += 1                                         # the reference's second loop reads a line before it looks at one

    def block(i):
        rows = []
        while i < len(lines):
            s = lines[i]
            i += 1
            if s and s[0] == "#":
                break
            if s:
                rows.append(s.split())
        return rows, i

    rows, i = block(i)
    out["hyper_priors"] = np.array([[float(r[0])] + [float(t) for t in r[2:]] for r in rows], dtype=np.float64) if rows else np.zeros((0, 1))
    out["hyper_priors_names"] = [r[1] for r in rows if len(r) > 1]       # io_ms_global.cpp:191-193
    rows, i = block(i)
    out["eigen_params"] = np.array([[float(t) for t in r] for r in rows], dtype=np.float64).reshape(-1, 6)
    rows, i = block(i)
    flat = [float(t) for r in rows for t in r]
    out["noise_params"] = np.array([-1.0] * (10 - len(flat)) + flat, dtype=np.float64)
    rows, i = block(i)
    s2 = np.full((10, 3), -1.0)
    if rows:
        s2[10 - len(rows):] = [[float(t) for t in r] for r in rows]
    out["noise_s2"] = s2
    rows, i = block(i)
    out["common_names"] = [r[0] for r in rows]
    out["common_names_priors"] = [r[1] for r in rows]
    mc = np.full((len(rows), 5), -9999.0)
    for k, r in enumerate(rows):
        vals = [float(t) for t in r[2:7]]
        mc[k, :len(vals)] = vals
    out["modes_common"] = mc
    for k in ("els",):
        out[k] = np.array(out[k], dtype=np.int64)
    out["freqs_ref"] = np.array(out["freqs_ref"], dtype=np.float64)
    return out


def read_local_model(path, slice_ind=0):
    """The local-fit `.model` dialect (read_MCMC_file_local, tamcmc/sources/io_local.cpp:25-327): the fields of MCMC_files for slice
    `slice_ind` of the file.  (The reader the reference has today indexes the second word of every hyper-prior row: a file whose
    'Extra parameters' block still holds its one-column rows -- like the shipped test/inputs/TF_3443483_local-v3.model -- makes it
    read out of bounds; such rows are rejected here.)"""
    out = read_ms_global_model(path, slice_ind=int(slice_ind))
    if out["freq_range"] is None:
        raise ValueError("slice %d: the file has fewer '*' frequency ranges" % slice_ind)
    if len(out["hyper_priors_names"]) != len(out["hyper_priors"]):
        raise ValueError("hyper-prior rows need a prior name in their second column (io_local.cpp:180: word[1])")
    return out


# ------------------------------------------------------------------------------------------------
# Custom tabulated priors, `<input root>_<k>.priors` beside the .model file (Config::setup, config.cpp:196-260; read with the
# same column reader as the spectra, read_data_ascii_Ncols): '#' comments, '!' labels, '*' units, then either two columns
# (x, PDF: a 1-D table for the `Tabulated(p)` prior) or a grid whose first row holds the x values after a corner token
# ('NA') and whose first column holds the y values (a 2-D table for `Tabulated_2d(p1,p2)`).
# ------------------------------------------------------------------------------------------------
def read_tabulated_prior(path):
    """-> dict(labels, units, ndim, and for ndim == 1: x, pdf; for ndim == 2: x (first row), y (first column), pdf[len(y), len(x)])"""
    labels, units, rows = [], [], []
    with open(path) as f:
        for line in f:
            s = line.strip()
            if not s or s[0] == "#":
                continue
            if s[0] == "!":
                labels = s[1:].split()
            elif s[0] == "*":
                units = s[1:].split()
            else:
                rows.append(s.split())
    if not rows:
        raise ValueError("no table in %s" % path)
    if len(rows[0]) == 2:
        a = np.array([[float(t) for t in r] for r in rows], dtype=np.float64)
        return {"labels": labels, "units": units, "ndim": 1, "x": a[:, 0].copy(), "pdf": a[:, 1].copy()}
    x = np.array([float(t) for t in rows[0][1:]], dtype=np.float64)
    body = np.array([[float(t) for t in r] for r in rows[1:]], dtype=np.float64)
    if body.shape[1] != x.size + 1:
        raise ValueError("2-D table: every row must hold the y value and one PDF value per x")
    return {"labels": labels, "units": units, "ndim": 2, "x": x, "y": body[:, 0].copy(), "pdf": body[:, 1:].copy()}


# ------------------------------------------------------------------------------------------------
# The other binary outputs of a run: statistical criteria, parallel-tempering log, proposal law, models
# (Outputs::write_bin_stat_criteria outputs.cpp:1472-1550, write_bin_parallel_temp_params :1336-1404,
#  write_bin_prop_params :1029-1229, write_bin_models :1406-1470).  Every file is a headerless stream of native little-endian
# values beside an ASCII `.hdr`; the stems are <dir><root><suffix> with the suffixes of the control file
# (config_default.cfg: _stat_criteria, _parallel_tempering, _proposals, _models).
# ------------------------------------------------------------------------------------------------
def stat_criteria_header_text(Nchains, Nsamples_done):
    labels = "".join("%s[%d]   " % (lab, i) for lab in ("logLikelihood", "logPrior", "logPosteriors") for i in range(Nchains))
    return ("# This is the header of the BINARY output file for the statistical information.\n"
            "# This file contains values for the logLikelihood (columns 0:Nchains-1), logPrior (columns Nchains:2*Nchains-1) and logPosterior (columns 2*Nchains:3*Nchains-1),  \n"
            "! Nsamples_done=%d\n! Nchains= %d\n! labels= %s\n" % (Nsamples_done, Nchains, labels))


def write_stat_criteria(stem, logL, logPrior, logPost, Nsamples_done=None, append=False, file_ext="bin"):
    """logL / logPrior / logPost: [Nsamples, Nchains].  Per sample the file holds the three rows one after the other."""
    logL, logPrior, logPost = (np.atleast_2d(np.asarray(a, dtype="<f8")) for a in (logL, logPrior, logPost))
    assert logL.shape == logPrior.shape == logPost.shape
    if not append:
        with open(stem + ".hdr", "w") as f:
            f.write(stat_criteria_header_text(logL.shape[1], logL.shape[0] if Nsamples_done is None else Nsamples_done))
    with open(stem + "." + file_ext, "ab" if append else "wb") as f:
        f.write(np.concatenate([logL, logPrior, logPost], axis=1).astype("<f8").tobytes())


def read_stat_criteria(stem, file_ext="bin"):
    """-> {'Nchains', 'Nsamples_done', 'logLikelihood' [N, Nchains], 'logPrior', 'logPosterior'}"""
    hdr = {}
    for l in open(stem + ".hdr"):
        if l.startswith("!") and "=" in l:
            k, v = l[1:].split("=", 1)
            hdr[k.strip()] = v.strip()
    nch = int(hdr["Nchains"])
    a = np.fromfile(stem + "." + file_ext, dtype="<f8").reshape(-1, 3 * nch)
    return {"Nchains": nch, "Nsamples_done": int(hdr["Nsamples_done"]), "logLikelihood": a[:, :nch].copy(), "logPrior": a[:, nch:2 * nch].copy(),
            "logPosterior": a[:, 2 * nch:].copy()}


# one record per sample: attempt_mixing (bool, 1 byte), chain0 (int32), Pswitch (double), switched (bool): 14 bytes, unpadded
PT_RECORD = np.dtype([("attempt_mixing", "?"), ("chain0", "<i4"), ("Pswitch", "<f8"), ("switched", "?")])


def parallel_tempering_header_text(Tcoefs, Nsamples_done):
    return ("# This is the header of the BINARY output file for the parameters of the parallel tempering.\n"
            "# This file contains values for \n"
            "# Correspondance between chain0=[0:Nchains-1] and temperature Tcoefs[chain] \n"
            "! Nsamples_done=%d\n! Tcoefs = %s\n! labels= attempt_mixing    chain0    Pswitch    switched \n" % (Nsamples_done, _eigen_row(np.asarray(Tcoefs, dtype=np.float64), False)))


def write_parallel_tempering(stem, Tcoefs, attempt_mixing, chain0, Pswitch, switched, Nsamples_done=None, append=False, file_ext="bin"):
    rec = np.zeros(len(chain0), dtype=PT_RECORD)
    rec["attempt_mixing"], rec["chain0"], rec["Pswitch"], rec["switched"] = attempt_mixing, chain0, Pswitch, switched
    if not append:
        with open(stem + ".hdr", "w") as f:
            f.write(parallel_tempering_header_text(Tcoefs, len(rec) if Nsamples_done is None else Nsamples_done))
    with open(stem + "." + file_ext, "ab" if append else "wb") as f:
        f.write(rec.tobytes())


def read_parallel_tempering(stem, file_ext="bin"):
    """-> (header dict with Tcoefs and Nsamples_done, structured array of PT_RECORD)"""
    hdr = {}
    for l in open(stem + ".hdr"):
        if l.startswith("!") and "=" in l:
            k, v = l[1:].split("=", 1)
            hdr[k.strip()] = v.strip()
    out = {"Nsamples_done": int(hdr["Nsamples_done"]), "Tcoefs": np.array([float(t) for t in hdr["Tcoefs"].split()])}
    return out, np.fromfile(stem + "." + file_ext, dtype=PT_RECORD)


def _proposal_header(what, Nchains, Nvars, Nsamples_done, var_names):
    first = "# This is the header of the BINARY output file for the parameters of the proposal law. These may vary if the MALA algorithm is learning.\n"
    if what == "sigmas":
        return first + "# This file contains only values for sigma[0:Nchains-1]\n! Nchains= %d\n! Nsamples_done=%d\n" % (Nchains, Nsamples_done)
    if what == "moves":
        return (first + "# This file contains only values for Pmove[0:Nchains-1] (first) and for moved[0:Nchains-1] (second group of Nchain values)\n"
                "! Nchains= %d\n! Nsamples_done=%d\n" % (Nchains, Nsamples_done))
    second = {"mus": "# This file contains only values for mu[0:Nchains-1][ 0:Nvars-1]. Each matrix is in a different file, indexed by the chain number\n",
              "covarmats": "# This file contains only values for covarmat[0:Nchains-1][ 0:Nvars-1][ 0:Nvars-1]. Each matrix is in a different file, indexed by the chain number\n"}[what]
    return first + second + "! Nchains= %d\n! Nvars= %d\n! Nsamples_done=%d\n! variable_names=%s\n" % (Nchains, Nvars, Nsamples_done, "".join(n + "   " for n in var_names))


def write_proposals(stem, sigmas, mus, covarmats, Pmoves, moveds, var_names, Nsamples_done=None, append=False, file_ext="bin"):
    """The proposal law over a run (Outputs::write_bin_prop_params): sigmas [N, Nchains]; mus [N, Nchains, Nvars]; covarmats
    [N, Nchains, Nvars, Nvars]; Pmoves [N, Nchains] doubles and moveds [N, Nchains] booleans, interleaved per sample.
    Files: <stem>_sigmas / _moves (one each), <stem>_mus_chain-k / _covarmats_chain-k (one per chain), four `.hdr`."""
    sigmas = np.atleast_2d(np.asarray(sigmas, dtype="<f8")); Pmoves = np.atleast_2d(np.asarray(Pmoves, dtype="<f8"))
    mus = np.asarray(mus, dtype="<f8"); covarmats = np.asarray(covarmats, dtype="<f8"); moveds = np.atleast_2d(np.asarray(moveds, dtype=bool))
    N, nch = sigmas.shape
    nv = mus.shape[2]
    assert mus.shape == (N, nch, nv) and covarmats.shape == (N, nch, nv, nv) and Pmoves.shape == (N, nch) and moveds.shape == (N, nch)
    done = N if Nsamples_done is None else Nsamples_done
    mode = "ab" if append else "wb"
    if not append:
        for what in ("sigmas", "moves", "mus", "covarmats"):
            with open("%s_%s.hdr" % (stem, what), "w") as f:
                f.write(_proposal_header(what, nch, nv, done, var_names))
    with open("%s_sigmas.%s" % (stem, file_ext), mode) as f:
        f.write(sigmas.tobytes())
    with open("%s_moves.%s" % (stem, file_ext), mode) as f:
        for i in range(N):
            f.write(Pmoves[i].tobytes() + moveds[i].astype("?").tobytes())
    for k in range(nch):
        with open("%s_mus_chain-%d.%s" % (stem, k, file_ext), mode) as f:
            f.write(np.ascontiguousarray(mus[:, k, :]).tobytes())
        with open("%s_covarmats_chain-%d.%s" % (stem, k, file_ext), mode) as f:
            f.write(np.ascontiguousarray(covarmats[:, k, :, :]).tobytes())          # row by row (ind_row outer, ind_col inner)


def read_proposals(stem, file_ext="bin"):
    hdr = {}
    for l in open(stem + "_mus.hdr"):
        if l.startswith("!") and "=" in l:
            k, v = l[1:].split("=", 1)
            hdr[k.strip()] = v.strip()
    nch, nv = int(hdr["Nchains"]), int(hdr["Nvars"])
    sig = np.fromfile("%s_sigmas.%s" % (stem, file_ext), dtype="<f8").reshape(-1, nch)
    mv = np.fromfile("%s_moves.%s" % (stem, file_ext), dtype=np.dtype([("Pmove", "<f8", (nch,)), ("moved", "?", (nch,))]))
    mus = np.stack([np.fromfile("%s_mus_chain-%d.%s" % (stem, k, file_ext), dtype="<f8").reshape(-1, nv) for k in range(nch)], axis=1)
    cov = np.stack([np.fromfile("%s_covarmats_chain-%d.%s" % (stem, k, file_ext), dtype="<f8").reshape(-1, nv, nv) for k in range(nch)], axis=1)
    return {"Nchains": nch, "Nvars": nv, "Nsamples_done": int(hdr["Nsamples_done"]), "variable_names": hdr["variable_names"].split(),
            "sigmas": sig, "Pmoves": mv["Pmove"].copy(), "moveds": mv["moved"].copy(), "mus": mus, "covarmats": cov}


def write_models(stem, models, append=False, file_ext="bin"):
    """models [N, Nchains, Ndata]: one file per chain, <stem>_models_chain-k (Outputs::write_bin_models).  (The reference opens
    the models' header with an empty file name, outputs.cpp:1419,1438: no `.hdr` is ever written for them.)"""
    models = np.asarray(models, dtype="<f8")
    for k in range(models.shape[1]):
        with open("%s_models_chain-%d.%s" % (stem, k, file_ext), "ab" if append else "wb") as f:
            f.write(np.ascontiguousarray(models[:, k, :]).tobytes())


def read_models(stem, Nchains, Ndata, file_ext="bin"):
    return np.stack([np.fromfile("%s_models_chain-%d.%s" % (stem, k, file_ext), dtype="<f8").reshape(-1, Ndata) for k in range(Nchains)], axis=1)
