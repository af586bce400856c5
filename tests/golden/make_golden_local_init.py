#!/usr/bin/env python
"""Golden answers for tamcmc-c_b200/model_setup.build_init_local from the reference's OWN read_MCMC_file_local + build_init_local
(tamcmc/sources/io_local.cpp:25-1176), compiled where they lie into oracle/_ref/libtamcmc_refio.so (make -C oracle refio).
Base text: the reference's test/inputs/TF_3443483_local-v3.model with the one-column rows of its obsolete 'Extra parameters' block taken
out (the reference's present reader indexes the second word of those rows and reads out of bounds on the shipped file); variants edit
the common-parameter block to reach the other branches.  Writes tests/golden/reference_local_init.json.  Needs /root/reference."""
import ctypes as C
import json
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtamcmc_refio.so"))
RESOL = 0.0123


def reference(text, slice_ind):
    with tempfile.NamedTemporaryFile("w", suffix=".model", delete=False) as f:
        f.write(text)
    cap = 1024
    n = C.c_int(0)
    inputs = np.zeros(cap); relax = np.zeros(cap, dtype=np.int32); priors = np.zeros((4, cap)); pl = np.zeros(11, dtype=np.int32); ex = np.zeros(10)
    names = C.create_string_buffer(cap * 64); pn = C.create_string_buffer(cap * 32); full = C.create_string_buffer(128)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = LIB.refio_build_init_local(f.name.encode(), int(slice_ind), C.c_double(RESOL), cap, C.byref(n), vp(inputs), vp(relax), vp(priors), vp(pl), vp(ex), names, pn, full)
    os.unlink(f.name)
    assert rc == 0
    N = n.value
    return {"model_fullname": full.value.decode(), "inputs": inputs[:N].tolist(), "relax": relax[:N].tolist(), "priors": priors[:, :N].tolist(),
            "plength": pl.tolist(), "extra_priors": ex.tolist(),
            "inputs_names": [names.raw[i * 64:(i + 1) * 64].split(b"\0")[0].decode() for i in range(N)],
            "priors_names": [pn.raw[i * 32:(i + 1) * 32].split(b"\0")[0].decode() for i in range(N)]}


def edit(text, drop=(), replace=None, add=()):
    out = []
    for l in text.splitlines():
        key = l.split()[0] if l.split() else ""
        if key in drop:
            continue
        if replace and key in replace:
            out.append(replace[key])
            continue
        out.append(l)
    return "\n".join(out + list(add)) + "\n"


def main():
    src = open("/root/reference/test/inputs/TF_3443483_local-v3.model").read().splitlines()
    base = "\n".join(l for l in src if l.strip() != "0.0000000") + "\n"
    cases = {}
    for s in range(8):
        cases["fixture_slice%d" % s] = (base, s)
    cases["hnlm"] = (edit(base, replace={"model_fullname": "           model_fullname              model_MS_local_Hnlm"}), 3)
    cases["hnlm_height_prior"] = (edit(base, replace={"model_fullname": "           model_fullname              model_MS_local_Hnlm",
                                                       "Height": "                   Height             Uniform          1.000000          5.0                 9000.00"}), 5)
    cases["amplitude_fix_auto"] = (edit(base, replace={"Height": "                Amplitude            Fix_Auto          10.000000          3.0"},
                                        add=["   fit_squareAmplitude_instead_Height    bool     1"]), 2)
    cases["height_fix_auto_width_user"] = (edit(base, replace={"Height": "                   Height            Fix_Auto          8.000000          2.5",
                                                                "Width": "                    Width             Uniform          0.500000          0.05          3.0"}), 4)
    cases["fixed_a1_inc_trunc"] = (edit(base, replace={"Splitting_a1": "             Splitting_a1                 Fix          0.700000",
                                                       "Inclination": "              Inclination                 Fix         95.000000"},
                                        add=["                  trunc_c                 Fix         25.000000"]), 1)
    cases["sqrt_keywords_asym_a3"] = (edit(base, drop=("Splitting_a1", "Inclination"),
                                           replace={"Splitting_a3": "             Splitting_a3             Uniform          0.010000         -0.1          0.1",
                                                    "Asymetry": "                 Asymetry            Jeffreys          10.000000          5.0          200.0",
                                                    "Asphericity_eta": "          Asphericity_eta             Uniform          1e-7          0.0          1e-5"},
                                           add=["  sqrt(splitting_a1).cosi             Uniform          0.500000          0.000000          1.200000",
                                                "  sqrt(splitting_a1).sini             Uniform          0.600000          0.000000          1.200000"]), 6)
    relaxed = []
    for l in base.splitlines():
        w = l.split()
        if w and w[0] == "p" and w[1] == "2":
            l = "p  2  %s  0      1       0" % w[2]
        if w and w[0] == "p" and w[1] == "3":
            l = "p  3  %s  1      0       1" % w[2]
        relaxed.append(l)
    cases["some_modes_fixed"] = ("\n".join(relaxed) + "\n", 3)
    out = {"resol": RESOL, "cases": {}}
    for name, (text, s) in cases.items():
        out["cases"][name] = {"model_text": text, "slice": s, "reference": reference(text, s)}
        print(name, out["cases"][name]["reference"]["plength"])
    json.dump(out, open(os.path.join(HERE, "reference_local_init.json"), "w"))


if __name__ == "__main__":
    main()
