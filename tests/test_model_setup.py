"""`.model` file -> parameter vector (tamcmc-c_b200/model_setup.py) against the reference's own build_init_MS_Global
(tamcmc/sources/io_ms_global.cpp:362-1400): golden answers written by tests/golden/make_golden_ms_global_init.py through
oracle/_ref/libtamcmc_refio.so (the reference's io_ms_global.cpp, io_models.cpp, string_handler.cpp compiled where they lie),
and -- where the reference tree and that library exist -- every MS_Global `.model` file the reference ships, live."""
import ctypes as C
import glob
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_ms_global_init.json")))
REFIO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libtamcmc_refio.so")
REF_INPUTS = "/root/reference/test/inputs"


def _compare(m, r):
    assert m["model_fullname"] == r["model_fullname"]
    assert list(m["plength"]) == list(r["plength"])
    ne = len(m["extra_priors"])          # 10 values (MS_Global) or 5 (red giants); the library's C entry pads its answer to 10
    assert np.array_equal(np.asarray(m["extra_priors"]), np.asarray(r["extra_priors"])[:ne]) and not np.any(np.asarray(r["extra_priors"])[ne:])
    assert list(m["inputs_names"]) == list(r["inputs_names"])
    assert list(m["priors_names"]) == list(r["priors_names"])
    assert np.array_equal(np.asarray(m["relax"]), np.asarray(r["relax"]))
    # values: the same IEEE operations in the same order -> bit for bit (sqrt / cos / sin of the a1, inclination pair included)
    assert np.array_equal(np.asarray(m["inputs"]), np.asarray(r["inputs"]))
    assert np.array_equal(np.asarray(m["priors"]), np.asarray(r["priors"]))


@pytest.mark.parametrize("case", sorted(GOLD["cases"]))
def test_build_init_ms_global_matches_reference_golden(pkg, tmp_path, case):
    c = GOLD["cases"][case]
    path = tmp_path / "case.model"
    path.write_text(c["model_text"])
    mf = pkg.formats.read_ms_global_model(str(path))
    build = pkg.model_setup.build_init_asymptotic if c["kind"] == "asymptotic" else pkg.model_setup.build_init_ms_global
    m = build(mf, GOLD["resol"])
    _compare(m, c["reference"])
    assert int(np.sum(m["plength"])) == len(m["inputs"])


def test_vector_layout_is_what_the_gpu_models_unpack(pkg, tmp_path):
    """The blocks of the flat vector sit where the expander reads them (expand.cu: heights | visibilities | l=0..3 frequencies |
    splittings | widths | noise | inclination | trunc_c, do_amp) and the model name maps to a GPU model id."""
    c = GOLD["cases"]["variant_classic"]
    path = tmp_path / "c.model"
    path.write_text(c["model_text"])
    m = pkg.model_setup.build_init_ms_global(pkg.formats.read_ms_global_model(str(path)), GOLD["resol"])
    pl = m["plength"]
    assert pkg.model_setup.GPU_MODEL_IDS[m["model_fullname"]] == 3
    assert pl[0] == pl[2] == pl[7] and pl[6] == 6 and pl[8] == 10 and pl[9] == 1 and pl[10] == 2
    o_split = int(pl[:6].sum())
    assert m["inputs_names"][o_split] == "Splitting_a1" and m["inputs_names"][o_split + 2] == "Splitting_a3" and m["inputs_names"][o_split + 5] == "Lorentzian_asymetry"
    o_cfg = int(pl[:10].sum())
    assert m["inputs"][o_cfg] == 30.0 and m["inputs"][o_cfg + 1] == 0.0
    fl0 = m["inputs"][int(pl[:2].sum()):int(pl[:3].sum())]
    assert np.all(np.diff(fl0) > 0)


def test_reference_exit_sites_raise(pkg, tmp_path):
    """Where the reference prints a diagnosis and exits, the restatement raises: an ajAlm model without filter_type
    (io_ms_global.cpp:1383-1386), an empty model name (:544-547), only one of the two sqrt(a1) keywords (:1180-1186)."""
    base = GOLD["cases"]["shipped_ajAlm_kplr003427720"]["model_text"]
    p = tmp_path / "a.model"
    p.write_text("\n".join(l for l in base.splitlines() if "filter_type" not in l) + "\n")
    with pytest.raises(ValueError, match="filter type"):
        pkg.model_setup.build_init_ms_global(pkg.formats.read_ms_global_model(str(p)), 0.01)
    p.write_text("\n".join(l for l in base.splitlines() if "model_fullname" not in l) + "\n")
    with pytest.raises(ValueError, match="Model name empty"):
        pkg.model_setup.build_init_ms_global(pkg.formats.read_ms_global_model(str(p)), 0.01)
    one = GOLD["cases"]["variant_a1etaa3_sqrt_keywords"]["model_text"]
    p.write_text("\n".join(l for l in one.splitlines() if ".sini" not in l) + "\n")
    with pytest.raises(ValueError, match="both sqrt"):
        pkg.model_setup.build_init_ms_global(pkg.formats.read_ms_global_model(str(p)), 0.01)


@pytest.mark.skipif(not (os.path.exists(REFIO) and os.path.isdir(REF_INPUTS)), reason="needs /root/reference and oracle/_ref/libtamcmc_refio.so")
def test_every_shipped_ms_global_model_file_live(pkg):
    lib = C.CDLL(REFIO)
    files = sorted(glob.glob(REF_INPUTS + "/Sun/*.model") + glob.glob(REF_INPUTS + "/Sun/fast/*.model") + glob.glob(REF_INPUTS + "/kplr*ajAlm*.model"))
    files = [f for f in files if not f.endswith("_Alm.model")]          # no filter_type line: the reference exits on it (tested above)
    assert len(files) >= 10
    for path in files:
        cap = 1024
        n = C.c_int(0)
        inputs = np.zeros(cap); relax = np.zeros(cap, dtype=np.int32); priors = np.zeros((4, cap)); pl = np.zeros(11, dtype=np.int32); ex = np.zeros(10)
        names = C.create_string_buffer(cap * 64); pn = C.create_string_buffer(cap * 32); full = C.create_string_buffer(128)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = lib.refio_build_init_ms_global(path.encode(), C.c_double(0.0123), cap, C.byref(n), vp(inputs), vp(relax), vp(priors), vp(pl), vp(ex), names, pn, full)
        assert rc == 0
        N = n.value
        r = {"model_fullname": full.value.decode(), "inputs": inputs[:N], "relax": relax[:N], "priors": priors[:, :N], "plength": pl, "extra_priors": ex,
             "inputs_names": [names.raw[i * 64:(i + 1) * 64].split(b"\0")[0].decode() for i in range(N)],
             "priors_names": [pn.raw[i * 32:(i + 1) * 32].split(b"\0")[0].decode() for i in range(N)]}
        m = pkg.model_setup.build_init_ms_global(pkg.formats.read_ms_global_model(path), 0.0123)
        _compare(m, r)


@pytest.mark.gpu
def test_model_file_to_gpu_loglikelihood(pkg, oracle):
    """End to end over the host side of the path: the `.model` text of a file the reference ships (Sun, aj model) -> parameter
    vector and plength (model_setup) -> model spectrum and log-likelihoods on the GPU (model id 23), against the CPU oracle on
    the same vector: 1e-10."""
    import tempfile
    c = GOLD["cases"]["shipped_aj_sun_fast"]
    with tempfile.NamedTemporaryFile("w", suffix=".model", delete=False) as f:
        f.write(c["model_text"])
    mf = pkg.formats.read_ms_global_model(f.name)
    os.unlink(f.name)
    m = pkg.model_setup.build_init_ms_global(mf, 0.0317)
    params, pl = np.ascontiguousarray(m["inputs"]), np.asarray(m["plength"], dtype=np.int32)
    model_id = pkg.model_setup.GPU_MODEL_IDS[m["model_fullname"]]
    assert model_id == 23
    lo, hi = mf["freq_range"]
    x = np.arange(lo, hi, 0.0317)
    rc, M = oracle.call_model(model_id, params, pl, x)[:2]
    assert rc == 0
    rng = np.random.default_rng(3)
    y = pkg.synth.chi2_2dof_spectrum(rng, M)
    relax = np.asarray(m["relax"]) != 0
    P = np.tile(params, (4, 1))
    P[1:, relax] *= 1.0 + 1e-3 * rng.standard_normal((3, int(relax.sum())))      # the sampler only moves the relaxed entries
    T = pkg.synth.tcoefs(4, 1.6)
    rc, L_ref = oracle.eval_chains(model_id, P, pl, x, y, T)
    assert rc == 0
    with pkg.Context(pkg.Star(model_id, pl, len(params), x, y), 4, T) as ctx:
        Mg = ctx.model(params)
        L, st = ctx.eval(ctx.pack_params([P]))
    assert (st == 0).all()
    assert np.max(np.abs(Mg - M) / np.abs(M)) < 1e-10
    assert np.max(np.abs(L[0] - L_ref) / np.abs(L_ref)) < 1e-10


@pytest.mark.gpu
def test_rgb_model_file_to_gpu_spectrum_against_the_reference(pkg):
    """The red-giant chain end to end: the text of the reference's fixture 10722175.model -> build_init_asymptotic (model_setup) ->
    ARMM host expander (tamcmc_host_expand_rgb_v4) -> GPU model spectrum, against what the REFERENCE's own chain (its
    build_init_asymptotic, then its model_RGB_asympt_aj_AppWidth_HarveyLike_v4) returned for the same file and frequency axis
    (tests/golden/reference_rgb_model_from_model_file.npz): 1e-10."""
    import tempfile
    G = np.load(os.path.join(HERE, "golden", "reference_rgb_model_from_model_file.npz"))
    V = np.load(os.path.join(HERE, "golden", "reference_rgb_vectors.npz"))
    x, y = V["x"], V["y"]
    with tempfile.NamedTemporaryFile("w", suffix=".model", delete=False) as f:
        f.write(GOLD["cases"]["rgb_shipped_10722175"]["model_text"])
    mf = pkg.formats.read_ms_global_model(f.name)
    os.unlink(f.name)
    m = pkg.model_setup.build_init_asymptotic(mf, float(G["resol"]))
    assert np.array_equal(m["inputs"], G["params"])
    model_id = pkg.model_setup.RGB_V4_MODELS[m["model_fullname"]]
    cap = 100
    row, nm = pkg.expand_rgb_v4(model_id, np.ascontiguousarray(m["inputs"]), np.asarray(m["plength"], dtype=np.int32), float(x[2] - x[1]), cap)
    mpl = pkg.synth.mode_table_plength(cap, int(m["plength"][8]), 1)
    with pkg.Context(pkg.Star(pkg.synth.MODEL_MODE_TABLE, mpl, len(row), x, y), 1, [1.0]) as ctx:
        M = ctx.model(row)
    assert np.max(np.abs(M - G["model"]) / np.abs(G["model"])) < 1e-10
