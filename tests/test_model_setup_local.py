"""The LOCAL-fit `.model` dialect: formats.read_local_model + model_setup.build_init_local against the reference's own
read_MCMC_file_local + build_init_local (tamcmc/sources/io_local.cpp:25-1176), through golden answers written by
tests/golden/make_golden_local_init.py from oracle/_ref/libtamcmc_refio.so (the reference's io_local.cpp, io_models.cpp, noise_models.cpp,
string_handler.cpp compiled where they lie) -- the eight slices of the reference's fixture TF_3443483 and seven variants that reach the
other branches (the Hnlm model, amplitudes, Fix_Auto heights, a user width prior, fixed a1 / inclination >= 90 / trunc_c, the sqrt(a1)
keywords with free a3 / asymmetry / asphericity, fixed modes) -- and live where the reference tree is present."""
import ctypes as C
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_local_init.json")))
REFIO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libtamcmc_refio.so")


def _compare(m, r):
    assert m["model_fullname"] == r["model_fullname"]
    assert list(m["plength"]) == list(r["plength"])
    assert np.array_equal(np.asarray(m["extra_priors"]), np.asarray(r["extra_priors"])[:4]) and not np.any(np.asarray(r["extra_priors"])[4:])
    assert list(m["inputs_names"]) == list(r["inputs_names"])
    assert list(m["priors_names"]) == list(r["priors_names"])
    assert np.array_equal(np.asarray(m["relax"]), np.asarray(r["relax"]))
    assert np.array_equal(np.asarray(m["inputs"]), np.asarray(r["inputs"]))          # same IEEE operations in the same order: bit for bit
    assert np.array_equal(np.asarray(m["priors"]), np.asarray(r["priors"]))


@pytest.mark.parametrize("case", sorted(GOLD["cases"]))
def test_build_init_local_matches_reference_golden(pkg, tmp_path, case):
    c = GOLD["cases"][case]
    path = tmp_path / "case.model"
    path.write_text(c["model_text"])
    m = pkg.model_setup.build_init_local(pkg.formats.read_local_model(str(path), c["slice"]), GOLD["resol"])
    _compare(m, c["reference"])
    assert int(np.sum(m["plength"])) == len(m["inputs"])
    assert pkg.model_setup.LOCAL_MODELS[m["model_fullname"]] in (11, 14)


def test_local_reader_and_exit_sites(pkg, tmp_path):
    base = GOLD["cases"]["fixture_slice3"]["model_text"]
    p = tmp_path / "a.model"
    p.write_text(base)
    mf = pkg.formats.read_local_model(str(p), 3)
    assert mf["freq_range"] == (127.56, 134.81) and mf["Dnu"] == 10.6795 and len(mf["els"]) == 21
    with pytest.raises(ValueError, match="fewer"):
        pkg.formats.read_local_model(str(p), 8)                                    # the file has eight slices
    # the shipped file itself: its obsolete one-column 'Extra parameters' rows are what the reference's reader trips over
    shipped = base.replace("# Extra parameters (obselete)\n", "# Extra parameters (obselete)\n0.0000000\n0.0000000\n")
    p.write_text(shipped)
    with pytest.raises(ValueError, match="prior name"):
        pkg.formats.read_local_model(str(p), 0)
    p.write_text("\n".join(l for l in base.splitlines() if "model_fullname" not in l) + "\n")
    with pytest.raises(ValueError, match="Model name empty"):
        pkg.model_setup.build_init_local(pkg.formats.read_local_model(str(p), 0), 0.01)
    one = GOLD["cases"]["sqrt_keywords_asym_a3"]["model_text"]
    p.write_text("\n".join(l for l in one.splitlines() if ".sini" not in l) + "\n")
    with pytest.raises(ValueError, match="both"):
        pkg.model_setup.build_init_local(pkg.formats.read_local_model(str(p), 6), 0.01)
    p.write_text(base.replace("* 127.56 134.81", "* 127.56 127.60"))              # a slice that holds no mode
    with pytest.raises(ValueError, match="No parameters"):
        pkg.model_setup.build_init_local(pkg.formats.read_local_model(str(p), 3), 0.01)


@pytest.mark.skipif(not (os.path.exists(REFIO) and os.path.isdir("/root/reference/tamcmc/sources")), reason="needs /root/reference and oracle/_ref/libtamcmc_refio.so")
def test_build_init_local_live_against_the_reference(pkg, tmp_path):
    lib = C.CDLL(REFIO)
    if not hasattr(lib, "refio_build_init_local"):
        pytest.skip("pin library built before io_local.cpp joined it")
    for name in ("fixture_slice5", "hnlm", "some_modes_fixed"):
        c = GOLD["cases"][name]
        path = tmp_path / (name + ".model")
        path.write_text(c["model_text"])
        for resol in (0.0317, 0.0079):
            cap = 1024
            n = C.c_int(0)
            inputs = np.zeros(cap); relax = np.zeros(cap, dtype=np.int32); priors = np.zeros((4, cap)); pl = np.zeros(11, dtype=np.int32); ex = np.zeros(10)
            names = C.create_string_buffer(cap * 64); pn = C.create_string_buffer(cap * 32); full = C.create_string_buffer(128)
            vp = lambda a: a.ctypes.data_as(C.c_void_p)
            rc = lib.refio_build_init_local(str(path).encode(), c["slice"], C.c_double(resol), cap, C.byref(n), vp(inputs), vp(relax), vp(priors), vp(pl), vp(ex), names, pn, full)
            assert rc == 0
            N = n.value
            r = {"model_fullname": full.value.decode(), "inputs": inputs[:N], "relax": relax[:N], "priors": priors[:, :N], "plength": pl, "extra_priors": ex,
                 "inputs_names": [names.raw[i * 64:(i + 1) * 64].split(b"\0")[0].decode() for i in range(N)],
                 "priors_names": [pn.raw[i * 32:(i + 1) * 32].split(b"\0")[0].decode() for i in range(N)]}
            _compare(pkg.model_setup.build_init_local(pkg.formats.read_local_model(str(path), c["slice"]), resol), r)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["fixture_slice3", "hnlm"])
def test_local_model_file_to_gpu_loglikelihood(pkg, oracle, tmp_path, case):
    """A slice of the local-fit `.model` text -> parameter vector and plength (model_setup.build_init_local) -> model spectrum and
    log-likelihoods on the GPU (model ids 11 / 14), against the CPU oracle on the same vector: 1e-10."""
    c = GOLD["cases"][case]
    path = tmp_path / "c.model"
    path.write_text(c["model_text"])
    mf = pkg.formats.read_local_model(str(path), c["slice"])
    m = pkg.model_setup.build_init_local(mf, 0.0317)
    params, pl = np.ascontiguousarray(m["inputs"]), np.asarray(m["plength"], dtype=np.int32)
    model_id = pkg.model_setup.LOCAL_MODELS[m["model_fullname"]]
    lo, hi = mf["freq_range"]
    x = np.arange(lo, hi, 0.0317)
    rc, M = oracle.call_model(model_id, params, pl, x)[:2]
    assert rc == 0
    rng = np.random.default_rng(5)
    y = pkg.synth.chi2_2dof_spectrum(rng, M)
    relax = np.asarray(m["relax"]) != 0
    P = np.tile(params, (4, 1))
    P[1:, relax] *= 1.0 + 1e-3 * rng.standard_normal((3, int(relax.sum())))
    T = pkg.synth.tcoefs(4, 1.6)
    rc, L_ref = oracle.eval_chains(model_id, P, pl, x, y, T)
    assert rc == 0
    with pkg.Context(pkg.Star(model_id, pl, len(params), x, y), 4, T) as ctx:
        Mg = ctx.model(params)
        L, st = ctx.eval(ctx.pack_params([P]))
    assert (st == 0).all()
    assert np.max(np.abs(Mg - M) / np.abs(M)) < 1e-10
    assert np.max(np.abs(L[0] - L_ref) / np.abs(L_ref)) < 1e-10
