#!/usr/bin/env python
"""Golden vectors of the reference's OWN build_init_MS_Global (tamcmc/sources/io_ms_global.cpp:362-1400) for
tests/test_model_setup.py: needs /root/reference and oracle/_ref/libtamcmc_refio.so (`make -C oracle refio`).

Cases: two of the `.model` files the reference ships (aj and ajAlm models), and variants of the first with the block of common
parameters rewritten for the other model families the function knows (Classic, Classic_v2, Classic_v3, a1l / a1n / a1nl etaa3,
a1etaa3 with and without the sqrt(a1) cos i / sin i keywords, amplitudes instead of heights with explicit frequency and width
priors and a numax line).  The text of every case's `.model` file is stored with the reference's answer, so the test needs
neither the reference tree nor the library.  The same for the red-giant dialect (build_init_asymptotic, io_asymptotic.cpp:32-875)
on the reference's fixture 10722175 and variants.  Writes tests/golden/reference_ms_global_init.json."""
import ctypes as C
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/test/inputs"
RESOL = 0.0317                                  # an arbitrary spectrum resolution (only the Fix_Auto width prior uses it)


def ref_build(lib, path, resol, fn="refio_build_init_ms_global"):
    cap = 1024
    n = C.c_int(0)
    inputs = np.zeros(cap); relax = np.zeros(cap, dtype=np.int32); priors = np.zeros((4, cap)); pl = np.zeros(11, dtype=np.int32); ex = np.zeros(10)
    names = C.create_string_buffer(cap * 64); pn = C.create_string_buffer(cap * 32); full = C.create_string_buffer(128)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = getattr(lib, fn)(path.encode(), C.c_double(resol), cap, C.byref(n), vp(inputs), vp(relax), vp(priors), vp(pl), vp(ex), names, pn, full)
    assert rc == 0, rc
    N = n.value
    return {"model_fullname": full.value.decode(), "inputs": inputs[:N].tolist(), "relax": relax[:N].tolist(), "priors": priors[:, :N].tolist(),
            "plength": pl.tolist(), "extra_priors": ex.tolist(),
            "inputs_names": [names.raw[i * 64:(i + 1) * 64].split(b"\0")[0].decode() for i in range(N)],
            "priors_names": [pn.raw[i * 32:(i + 1) * 32].split(b"\0")[0].decode() for i in range(N)]}


COMMON_TAIL = """            Visibility_l1            Gaussian          1.500000          1.500000          0.100000
            Visibility_l2            Gaussian          0.530000          0.530000          0.030000
            Visibility_l3            Gaussian          0.080000          0.080000          0.020000
                   Height            Jeffreys          1.000000          1000.000
                    Width            Fix_Auto          1.000000
                  trunc_c                 Fix          30.0000
"""
A1_INC = """             Splitting_a1             Uniform          0.900000          0.000000          4.00000
             Splitting_a3             Uniform          0.010000         -0.200000          0.20000
                 Asymetry        Jeffreys_abs         10.000000          5.000000          200.0000
              Inclination             Uniform          62.50000          0.000000          90.0000
"""
VARIANTS = {
    "classic": "model_fullname model_MS_Global_a1etaa3_HarveyLike_Classic\nfreq_smoothness bool 1.0 2.0\n" + A1_INC + COMMON_TAIL,
    "classic_v2": "model_fullname model_MS_Global_a1etaa3_HarveyLike_Classic_v2\nfreq_smoothness bool 0.0 1.5\n" + A1_INC + COMMON_TAIL,
    "classic_v3": "model_fullname model_MS_Global_a1etaa3_HarveyLike_Classic_v3\n" + A1_INC + COMMON_TAIL,
    "a1l": "model_fullname model_MS_Global_a1l_etaa3_HarveyLike\n" + A1_INC + COMMON_TAIL,
    "a1n": "model_fullname model_MS_Global_a1n_etaa3_HarveyLike\n" + A1_INC + COMMON_TAIL,
    "a1nl": "model_fullname model_MS_Global_a1nl_etaa3_HarveyLike\n" + A1_INC + COMMON_TAIL,
    "a1etaa3_from_a1_inc": "model_fullname model_MS_Global_a1etaa3_HarveyLike\n" + A1_INC + COMMON_TAIL,
    "a1etaa3_fixed_a1_inc": "model_fullname model_MS_Global_a1etaa3_HarveyLike\n" + A1_INC.replace("Uniform          0.900000          0.000000          4.00000", "Fix              0.900000")
                            .replace("Uniform          62.50000          0.000000          90.0000", "Fix          62.50000") + COMMON_TAIL,
    "a1etaa3_sqrt_keywords": "model_fullname model_MS_Global_a1etaa3_HarveyLike\n" + A1_INC +
                             "sqrt(splitting_a1).cosi Uniform 0.45 0.0 2.0\nsqrt(splitting_a1).sini Uniform 0.80 0.0 2.0\n" + COMMON_TAIL,
    "classic_amplitudes": "model_fullname model_MS_Global_a1etaa3_HarveyLike_Classic\nfit_squareAmplitude_instead_Height bool 1\n" + A1_INC +
                          COMMON_TAIL.replace("Fix_Auto          1.000000", "Jeffreys          0.050000          25.0000").replace("Fix          30.0000", "Fix          -1.0") +
                          "Frequency GUG -1 -1 -1 0.5 0.7\n",
}


def variant_text(base_text, common, numax_line=None):
    """the base file with everything behind its '# Controls and priors for common parameters' line replaced"""
    lines = base_text.splitlines()
    k = [i for i, l in enumerate(lines) if l.strip().startswith("#") and "common parameters" in l][0]
    head = lines[:k + 1]
    if numax_line:
        j = [i for i, l in enumerate(head) if l.strip().startswith("!") and not l.strip().startswith("!!") and not l.strip().startswith("!n")][0]
        head.insert(j + 1, numax_line)
    return "\n".join(head) + "\n" + common


def main():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtamcmc_refio.so"))
    base = open(os.path.join(REF, "Sun", "fast", "19992002_incfix_fast_Priorevalrange.model")).read()
    cases = {"shipped_aj_sun_fast": base,
             "shipped_ajAlm_kplr003427720": open(os.path.join(REF, "kplr003427720_kasoc-psd_slc_v1_ajAlm_gate.model")).read()}
    for name, common in VARIANTS.items():
        cases["variant_" + name] = variant_text(base, common, "!n 3000.0 120.0" if name == "classic_amplitudes" else None)
    # the red-giant dialect (build_init_asymptotic, io_asymptotic.cpp:32-875): the reference's fixture 10722175 with and without the
    # bias spline, the constant-width model, amplitudes + an explicit frequency prior
    rgb = open(os.path.join(REF, "RGB", "v1.86.0", "10722175.model")).read()
    cases["rgb_shipped_10722175"] = rgb
    cases["rgb_shipped_10722175_nobias"] = open(os.path.join(REF, "RGB", "v1.86.0", "10722175_nobias.model")).read()
    cases["rgb_variant_ctewidth"] = rgb.replace("model_RGB_asympt_aj_AppWidth_HarveyLike_v4", "model_RGB_asympt_aj_CteWidth_HarveyLike_v4")
    cases["rgb_variant_amplitudes_frequency"] = rgb.rstrip("\n") + "\n fit_squareAmplitude_instead_Height bool 1\n Frequency GUG -1 -1 -1 0.05 0.07\n"
    out = {"resol": RESOL, "generator": "tests/golden/make_golden_ms_global_init.py", "cases": {}}
    for name, text in cases.items():
        with tempfile.NamedTemporaryFile("w", suffix=".model", delete=False) as f:
            f.write(text)
        sys.stdout.flush()
        kind = "asymptotic" if name.startswith("rgb_") else "ms_global"
        r = ref_build(lib, f.name, RESOL, "refio_build_init_" + kind)
        os.unlink(f.name)
        out["cases"][name] = {"kind": kind, "model_text": text, "reference": r}
        print("%-36s %-52s N = %d" % (name, r["model_fullname"], len(r["inputs"])), file=sys.stderr)
    with open(os.path.join(HERE, "reference_ms_global_init.json"), "w") as f:
        json.dump(out, f)
    # the spectrum the reference's OWN model function (model_RGB_asympt_aj_AppWidth_HarveyLike_v4, models.cpp:4684-5079) returns for
    # the vector its OWN build_init_asymptotic makes of the fixture's .model file, on the frequency axis of reference_rgb_vectors.npz
    # (the fixture's 80-128 microHz slice): the GPU test goes .model text -> model_setup -> host expander -> GPU and lands on it
    os.environ["OMP_NUM_THREADS"] = "1"
    sys.path.insert(0, os.path.dirname(HERE))
    import _refshim
    R = _refshim.get()
    x = np.load(os.path.join(HERE, "reference_rgb_vectors.npz"))["x"]
    step = float(x[2] - x[1])
    with tempfile.NamedTemporaryFile("w", suffix=".model", delete=False) as f:
        f.write(cases["rgb_shipped_10722175"])
    r = ref_build(lib, f.name, step, "refio_build_init_asymptotic")
    os.unlink(f.name)
    rc, M = R.call_model(25, np.array(r["inputs"]), np.array(r["plength"], dtype=np.int32), x)
    assert rc == 0 and np.all(np.isfinite(M))
    np.savez_compressed(os.path.join(HERE, "reference_rgb_model_from_model_file.npz"), model=M, resol=step, params=np.array(r["inputs"]))
    print("rgb model from the .model file: range %.4g..%.4g" % (M.min(), M.max()), file=sys.stderr)


if __name__ == "__main__":
    main()
