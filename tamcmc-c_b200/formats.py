"""Readers for the reference's simplest on-disk inputs (first slice): the ASCII `.data` spectrum and
the "simple matrix" `.model` file of the Gaussian-envelope fits (model_Harvey_Gaussian / model_Kallinger2014_Gaussian).

  .data  : '#' comments, '!' column labels, '*' units, then whitespace-separated columns (frequency, power[, ...])
           -- Config::read_data_ascii_Ncols (tamcmc/sources/config.cpp)
  .model : '#' comments, '* fmin fmax' the fitted range, '! names', the initial values, '! relax' + one 0/1 flag per
           parameter, '! prior names', then up to four rows of prior parameters (-9999 = unused slot)
           -- Config::read_inputs_prior_Simple_Matrix (config.cpp:560-660)

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
