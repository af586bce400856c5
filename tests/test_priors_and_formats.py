"""Either side of the hot path for the simplest model family (first slice): the reference's generic priors
restated for the C++ driver (tamcmc-c_b200/host/priors.hpp) against values from the REFERENCE's own stats_dictionary.cpp /
priors_calc.cpp, the readers of its `.data` / simple-matrix `.model` files, and its real fixture 10280410_Gaussfit run on the
GPU (tests/golden/make_golden_priors_and_gaussfit.py generated the vectors)."""
import json
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD_P = os.path.join(HERE, "golden", "reference_priors.json")
GOLD_G = os.path.join(HERE, "golden", "reference_gaussfit_10280410.npz")


def _build():
    exe = os.path.join(HERE, "cpp", "test_priors")
    src = os.path.join(HERE, "cpp", "test_priors.cpp")
    hdr = os.path.join(ROOT, "tamcmc-c_b200", "host", "priors.hpp")
    if not (os.path.exists(exe) and os.path.getmtime(exe) > max(os.path.getmtime(src), os.path.getmtime(hdr))):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-o", exe, src])
    return exe


def _same(a, b):
    if np.isinf(b) or np.isnan(b):
        return (np.isinf(a) and np.isinf(b) and (a < 0) == (b < 0)) or (np.isnan(a) and np.isnan(b))
    return abs(a - b) <= 1e-13 * max(1.0, abs(b))


def test_generic_priors_match_reference():
    g = json.load(open(GOLD_P))
    lines, want = [], []
    for kind, a, b, c, d, x, v in g["primitive"]:
        lines.append("P %d %r %r %r %r %r" % (kind, a, b, c, d, x)); want.append(v)
    for s in g["sets"]:
        tag = {-1: "G", 1: "H", 0: "K"}[s["which"]]
        flat = " ".join(repr(float(v)) for row in s["pri"] for v in row)
        lines.append("%s %d %s %s %s" % (tag, len(s["params"]), " ".join(repr(float(v)) for v in s["params"]),
                                         " ".join(str(int(k)) for k in s["kinds"]), flat))
        want.append(s["value"])
    r = subprocess.run([_build()], input="\n".join(lines) + "\n", stdout=subprocess.PIPE, text=True, check=True)
    got = [float(t) for t in r.stdout.split()]
    assert len(got) == len(want)
    bad = [(l, a, b) for l, a, b in zip(lines, got, want) if not _same(a, b)]
    assert not bad, bad[:3]
    assert sum(1 for v in want if np.isinf(v)) > 10 and sum(1 for v in want if np.isfinite(v)) > 100     # both branches exercised


def test_simple_matrix_model_and_data_readers(pkg, tmp_path):
    g = np.load(GOLD_G)
    f = tmp_path / "star.model"
    f.write_text(str(g["model_text"]))
    m = pkg.formats.read_simple_matrix_model(str(f))
    assert m["names"] == ["H1", "tc1", "p1", "H2", "tc2", "p2", "B0", "Amax", "numax", "Gauss_sigma"]
    assert np.array_equal(m["inputs"], g["rows"][0])
    assert list(m["relax"]) == [1, 1, 0, 1, 1, 1, 1, 1, 1, 1]
    assert m["prior_names"][2] == "Fix" and m["prior_kinds"][2] == 0 and m["prior_kinds"][9] == 7        # GUG on the envelope width
    assert m["priors"].shape == (4, 10) and m["priors"][2, 0] == -9999.0 and m["priors"][3, 9] == 42.544956
    assert m["xrange"] == [0.000079, 256.875763]
    d = tmp_path / "star.data"
    d.write_text("# comment\n! frequency power\n* (microHz) (ppm^2/microHz)\n 1.0 5.0\n 2.0 6.0\n 3.5 7.0\n")
    x, y = pkg.formats.read_data(str(d), xrange=(1.5, 10))
    assert list(x) == [2.0, 3.5] and list(y) == [6.0, 7.0]


def test_oracle_on_reference_gaussfit_fixture(pkg, oracle):
    g = np.load(GOLD_G)
    for r, M_ref, L_ref in zip(g["rows"], g["model"], g["logL"]):
        rc, M = oracle.call_model(1, r, pkg.synth.ENVELOPE_PLENGTH, g["x"])
        assert rc == 0 and np.max(np.abs(M - M_ref) / M_ref) < 1e-13
        assert abs(oracle.chi22p(g["y"], M, 1) - L_ref) <= 1e-12 * abs(L_ref)


@pytest.mark.gpu
def test_gpu_on_reference_gaussfit_fixture(pkg):
    """Real Kepler spectrum of KIC 10280410 + the reference's own .model initial values -> GPU model and logL against the
    REFERENCE's model_Harvey_Gaussian / likelihood_chi22p."""
    g = np.load(GOLD_G)
    x, y, rows = g["x"], g["y"], g["rows"]
    T = pkg.synth.tcoefs(len(rows), 1.7)
    with pkg.Context(pkg.Star(1, pkg.synth.ENVELOPE_PLENGTH, rows.shape[1], x, y), len(rows), T) as ctx:
        for r, M_ref in zip(rows, g["model"]):
            assert np.max(np.abs(ctx.model(r) - M_ref) / M_ref) < 1e-10
        L, st = ctx.eval(rows)
        assert (st == 0).all()
        assert np.max(np.abs(L[0] * T - g["logL"]) / np.abs(g["logL"])) < 1e-10


def _build_gaussfit():
    exe = os.path.join(HERE, "cpp", "test_gaussfit")
    src = os.path.join(HERE, "cpp", "test_gaussfit.cpp")
    libdir = os.path.join(ROOT, "tamcmc-c_b200")
    deps = [src, os.path.join(libdir, "host", "priors.hpp"), os.path.join(libdir, "host", "mcmc_driver.hpp"), os.path.join(ROOT, "include", "tamcmc_gpu.h")]
    if not (os.path.exists(exe) and os.path.getmtime(exe) > max(os.path.getmtime(d) for d in deps)):
        cuda_lib = "/usr/local/cuda/lib64"
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-fopenmp", "-o", exe, src, "-L" + libdir, "-ltamcmc_gpu", "-L" + cuda_lib, "-lcudart",
                               "-Wl,-rpath," + libdir, "-Wl,-rpath," + cuda_lib])
    return exe


def test_gaussfit_example_builds(pkg):
    pkg.lib()
    assert os.path.exists(_build_gaussfit())


@pytest.mark.gpu
def test_gaussfit_mcmc_on_reference_fixture(pkg, tmp_path):
    """The whole first-stage analysis of the reference on its own fixture: read .model/.data, sample the posterior of the
    Harvey + Gaussian model under the .model's priors with the GPU likelihood.  The chain leaves the .model's rough guess
    (log posterior +5000), stays inside the priors, is reproducible for a fixed seed, and mixes (measured: 28k MCMC steps/s
    of 6 chains x 32k bins on one B200; 6000 and 30000 steps give the same posterior means)."""
    g = np.load(GOLD_G)
    f = tmp_path / "star.model"
    f.write_text(str(g["model_text"]))
    m = pkg.formats.read_simple_matrix_model(str(f))
    x, y = g["x"], g["y"]
    errors = 0.03 * np.abs(m["inputs"])
    case = tmp_path / "case.bin"
    with open(case, "wb") as fh:
        for a in ([len(x), len(m["inputs"]), 6, 20261018], x, y, m["inputs"], m["relax"], m["prior_kinds"], m["priors"].ravel(), errors):
            fh.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    exe = _build_gaussfit()
    outs = []
    for _ in range(2):
        r = subprocess.run([exe, str(case), "6000"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout
        outs.append(json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1]))
    a, b = outs
    assert a["final"] == b["final"]                                  # same seed, same chain
    assert a["mean"][2] == 4.0 and a["sd"][2] == 0.0                 # p1 is fixed in the .model (relax 0)
    lo, hi = m["priors"][0, 8], m["priors"][1, 8]                    # uniform prior on numax
    assert lo < a["mean"][8] < hi and a["sd"][8] < 15.0
    assert a["mean"][9] > 0.5 * 0.263 * a["mean"][8] ** 0.77         # the Stello+2009 width condition of priors_Harvey_Gaussian held
    assert a["logpost_final"] > a["logpost_initial"] + 1000.0        # the .model's initial guess is far from the posterior mode
    assert 0.03 < a["acceptance_cold"] < 0.8 and a["swap_rate"] > 0.05
    print(a)


def test_params_output_header_matches_the_reference_byte_for_byte(pkg):
    """Outputs::write_bin_params (outputs.cpp:1231-1334): the header the reference wrote for its own Gaussian-envelope run
    (tests/golden/reference_params_hdr.json, made by make_golden_params_hdr.py) is parsed to its values, and the writer
    regenerates the same bytes from them."""
    fmt = pkg.formats
    gold = json.load(open(os.path.join(HERE, "golden", "reference_params_hdr.json")))
    h = fmt.parse_params_header(gold["text"])
    tok = gold["tokens"]
    for k in ("Nsamples", "Nchains", "Nsamples_done", "Nvars", "Ncons"):
        assert h[k] == int(tok[k][0])
    assert h["relax"].tolist() == [int(t) for t in tok["relax"]] and h["plength"].tolist() == [int(t) for t in tok["plength"]]
    assert h["constant_names"] == tok["constant_names"] and h["variable_names"] == tok["variable_names"]
    assert h["constant_values"].tolist() == [float(t) for t in tok["constant_values"]]
    assert len(h["variable_names"]) == h["Nvars"] and len(h["relax"]) == h["Nvars"] + h["Ncons"]
    text = fmt.params_header_text(h["Nsamples"], h["Nchains"], h["Nsamples_done"], h["relax"], h["plength"], h["constant_names"],
                                  h["constant_values"], h["variable_names"])
    assert text == gold["text"]
    # Eigen pads every coefficient of a printed row to the widest one
    assert fmt._eigen_row([20, 3, 20, 6], True) == "20  3 20  6" and fmt._eigen_row([0.5, 12.25], False) == "  0.5 12.25"
    assert "! constant_values= -1\n" in fmt.params_header_text(10, 1, 10, [1], [1], ["None"], [], ["a"])


def test_params_outputs_round_trip(pkg, tmp_path):
    fmt = pkg.formats
    rng = np.random.default_rng(0)
    names = ["H1", "tc1", "B0"]
    s1, s2 = rng.normal(size=(7, 3, 3)), rng.normal(size=(5, 3, 3))
    prefix = str(tmp_path / "star_A_params")
    fmt.write_params_outputs(prefix, s1, [1, 1, 0, 1], [1, 1, 1, 1], ["p1"], [4.0], names, Nsamples=12)
    fmt.write_params_outputs(prefix, s2, [1, 1, 0, 1], [1, 1, 1, 1], ["p1"], [4.0], names, append=True)      # second buffer
    h, got = fmt.read_params_outputs(prefix)
    assert h["Nsamples"] == 12 and h["Nchains"] == 3 and h["Nvars"] == 3 and h["variable_names"] == names
    assert got.shape == (12, 3, 3) and np.array_equal(got, np.concatenate([s1, s2]))         # bit-exact: raw float64
    assert os.path.getsize(prefix + "_chain-2.bin") == 12 * 3 * 8
    _, one = fmt.read_params_outputs(prefix, chains=[1])
    assert np.array_equal(one[:, 0, :], got[:, 1, :])
    with pytest.raises(ValueError):
        fmt.write_params_outputs(prefix, s1[:, :, :2], [1], [1], ["None"], [], names)


def test_cpp_outputs_writer_matches_reference_header_and_python_reader(pkg, tmp_path):
    """host/outputs.hpp (the C++ driver side of Outputs::write_bin_params): same header bytes as the reference's fixture, chain
    files the Python reader takes back bit for bit."""
    exe, src = os.path.join(HERE, "cpp", "test_outputs"), os.path.join(HERE, "cpp", "test_outputs.cpp")
    hdr = os.path.join(HERE, "..", "tamcmc-c_b200", "host", "outputs.hpp")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-o", exe, src])
    prefix = str(tmp_path / "10280410_Gaussfit_A_params")
    r = subprocess.run([exe, prefix], stdout=subprocess.PIPE, text=True, check=True)
    assert r.stdout == "  0.5 12.25"
    gold = json.load(open(os.path.join(HERE, "golden", "reference_params_hdr.json")))
    assert open(prefix + ".hdr").read() == gold["text"]
    h, s = pkg.formats.read_params_outputs(prefix)
    assert s.shape == (10, 4, 9)
    i, c, v = np.meshgrid(np.arange(10), np.arange(4), np.arange(9), indexing="ij")
    assert np.array_equal(s, 1000.0 * i + 10.0 * c + v + 0.125)


def test_acceptance_log_and_restore_files_of_the_reference(pkg, tmp_path):
    """The acceptance log regenerates byte for byte from its parsed values (Outputs::write_txt_acceptance, outputs.cpp:747-789);
    the three restore files the reference wrote (write_buffer_restore, outputs.cpp:863-1027) parse into the state a restart
    needs.  Texts: tests/golden/reference_outputs_10280410.json (make_golden_params_hdr.py)."""
    fmt = pkg.formats
    gold = json.load(open(os.path.join(HERE, "golden", "reference_outputs_10280410.json")))
    p = tmp_path / "acc.txt"
    p.write_text(gold["acceptance"])
    x, r = fmt.read_acceptance(str(p))
    assert r.shape[1] == 4 and x[0] == 2500 and r[0].tolist() == [0.237, 0.2316, 0.2764, 0.2348] and np.all(np.diff(x) == 5000)
    assert fmt.acceptance_text(x, r) == gold["acceptance"]
    for n in (1, 2, 3):
        (tmp_path / ("10280410_Gaussfit_restore_A_%d.dat" % n)).write_text(gold["restore"][str(n)])
    st = fmt.read_restore(str(tmp_path), "10280410_Gaussfit")
    assert st["Nchains"] == 4 and st["Nvars"] == 9 and st["iteration"] == 99999 and st["variable_names"][-1] == "Gauss_sigma"
    for k in ("vars", "vars_mean", "mus", "mus_mean"):
        assert st[k].shape == (4, 9)
    assert st["vars"][0, 0] == 1110.3 and st["vars"][3, 8] == 11.8951 and st["mus"][3, 2] == 345.202
    assert st["sigmas"].tolist() == [0.00234744, 0.00110771, 0.0414778, 0.00270788] and np.array_equal(st["sigmas"], st["sigmas_mean"])
    for k in ("covarmats", "covarmats_mean"):
        C = st[k]
        assert C.shape == (4, 9, 9)
        assert np.allclose(C, np.transpose(C, (0, 2, 1)), rtol=1e-5, atol=0)          # symmetric to the printed digits
        assert np.all(np.linalg.eigvalsh(0.5 * (C + np.transpose(C, (0, 2, 1)))) > 0)       # proposal covariances
    assert st["covarmats"][0, 0, 0] == 4.20452 and st["covarmats"][0, 1, 0] == 1.11062


def test_restore_writer_reproduces_the_reference_data_lines(pkg, tmp_path):
    """write_restore: every non-comment line of the three files the reference wrote comes back byte for byte from the parsed
    state (the comment lines of file 1 differ between versions of the reference: `do_restore=1` vs `do_restore_[X]=1`)."""
    fmt = pkg.formats
    gold = json.load(open(os.path.join(HERE, "golden", "reference_outputs_10280410.json")))
    for n in (1, 2, 3):
        (tmp_path / ("10280410_Gaussfit_restore_A_%d.dat" % n)).write_text(gold["restore"][str(n)])
    st = fmt.read_restore(str(tmp_path), "10280410_Gaussfit")
    texts = fmt.restore_texts(st)
    data = lambda t: [l for l in t.splitlines() if not l.startswith("#")]
    for n in (1, 2, 3):
        assert data(texts[n]) == data(gold["restore"][str(n)]), n
        assert len([l for l in texts[n].splitlines() if l.startswith("#")]) == len([l for l in gold["restore"][str(n)].splitlines() if l.startswith("#")])
    assert texts[3] == gold["restore"]["3"] and texts[2] == gold["restore"]["2"]
    out = tmp_path / "again"
    out.mkdir()
    fmt.write_restore(str(out), "star", st, phase="L")
    st2 = fmt.read_restore(str(out), "star", phase="L")
    assert all(np.array_equal(st[k], st2[k]) for k in ("vars", "vars_mean", "sigmas", "mus", "covarmats", "covarmats_mean"))


def test_cpp_restore_reader_agrees_with_python(pkg, tmp_path):
    """host/outputs.hpp:read_restore on the reference's own restore files: same state as formats.read_restore, value for value."""
    exe, src = os.path.join(HERE, "cpp", "test_outputs"), os.path.join(HERE, "cpp", "test_outputs.cpp")
    hdr = os.path.join(HERE, "..", "tamcmc-c_b200", "host", "outputs.hpp")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-o", exe, src])
    gold = json.load(open(os.path.join(HERE, "golden", "reference_outputs_10280410.json")))
    for n in (1, 2, 3):
        (tmp_path / ("10280410_Gaussfit_restore_A_%d.dat" % n)).write_text(gold["restore"][str(n)])
    st = pkg.formats.read_restore(str(tmp_path), "10280410_Gaussfit")
    r = subprocess.run([exe, "restore", str(tmp_path), "10280410_Gaussfit", "A"], stdout=subprocess.PIPE, text=True, check=True)
    lines = r.stdout.strip().splitlines()
    assert lines[0] == "rc 0 Nchains 4 Nvars 9 iteration 99999 names 9 last Gauss_sigma"
    for line, key in zip(lines[1:], ("vars", "vars_mean", "sigmas", "sigmas_mean", "mus", "mus_mean", "covarmats", "covarmats_mean")):
        vals = np.array([float(t) for t in line.split()[1:]])
        assert int(line.split()[0]) == st[key].size and np.array_equal(vals, st[key].ravel()), key
    assert subprocess.run([exe, "restore", str(tmp_path), "missing", "A"], stdout=subprocess.PIPE).returncode == 1
    # write_restore (C++): files 2 and 3 come back byte for byte, file 1 on every data line (as with the Python writer)
    out = tmp_path / "again"
    out.mkdir()
    subprocess.run([exe, "restore", str(tmp_path), "10280410_Gaussfit", "A", str(out)], stdout=subprocess.PIPE, check=True)
    data = lambda t: [l for l in t.splitlines() if not l.startswith("#")]
    for n in (1, 2, 3):
        again = (out / ("10280410_Gaussfit_restore_A_%d.dat" % n)).read_text()
        assert data(again) == data(gold["restore"][str(n)]), n
        if n > 1:
            assert again == gold["restore"][str(n)]
        assert again == pkg.formats.restore_texts(st)[n]


def test_cfg_reader(pkg, tmp_path):
    """read_cfg / mala_config follow Config::format_line and read_cfg_file (config.cpp:1062-1110, 1223-1300): value up to the
    first ';', strtod-style numbers, comma lists, '#' comment lines, '!Group:' headers.  The values of the reference's shipped
    config_default.cfg as read by the same code are in tests/golden/reference_cfg_default.json (make_golden_cfg.py)."""
    fmt = pkg.formats
    gold = json.load(open(os.path.join(HERE, "golden", "reference_cfg_default.json")))
    assert gold["MALA"] == {"Nchains": 5, "lambda_temp": 3.5, "c0": 10.0, "epsilon1": 1e-12, "epsi2": 1e-12, "A1": 1e14,
                            "target_acceptance": 0.234, "dN_mixing": 1, "Nt_learn": [1000, 1500, 100000], "periods_learn": [1, 1]}
    assert gold["Modeling"]["likelihood_fct_name"] == "chi(2,2p)" and gold["groups"] == ["Data", "Diagnostics", "MALA", "Modeling", "Outputs"]
    sample = """# a control file in the reference's syntax
!MALA:
\ttarget_acceptance=0.234;
\tc0=10;  Relaxation constrain. Important: adjust; empirically
\tepsilon1=1e-12;  free text
\tepsilon2=1e-10;
\tA1=1e14;
\tNt_learn=200, 400, 5000;   text
\tperiods_learn=1, 2;
 \t#Nchains=10;   commented out
\tNchains=4;
\tdN_mixing=1.;  trailing dot
 \tlambda_temp=1.70 #3.50; the number ends where strtod stops
!Data:
\tysig_col=-1 //-1;  // text
"""
    p = tmp_path / "c.cfg"
    p.write_text(sample)
    g = fmt.read_cfg(str(p))
    m = fmt.mala_config(g)
    assert m == {"Nchains": 4, "lambda_temp": 1.7, "c0": 10.0, "epsilon1": 1e-12, "epsi2": 1e-10, "A1": 1e14, "target_acceptance": 0.234,
                 "dN_mixing": 1, "Nt_learn": [200, 400, 5000], "periods_learn": [1, 2]}
    assert fmt.cfg_number(g["Data"]["ysig_col"]) == -1
    # the C++ reader of the driver side (host/config.hpp) on the same file
    exe, src = os.path.join(HERE, "cpp", "test_config"), os.path.join(HERE, "cpp", "test_config.cpp")
    hdrs = [os.path.join(HERE, "..", "tamcmc-c_b200", "host", h) for h in ("config.hpp", "mcmc_driver.hpp")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in hdrs]):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-fopenmp", "-o", exe, src])
    r = subprocess.run([exe, str(p)], stdout=subprocess.PIPE, text=True, check=True)
    t = r.stdout.split()
    got = {t[i]: t[i + 1] for i in range(0, t.index("Nt_learn"), 2)}
    assert got["ok"] == "1" and got["groups"] == "2" and int(got["Nchains"]) == m["Nchains"] and int(got["dN_mixing"]) == m["dN_mixing"]
    for k in ("lambda_temp", "c0", "epsilon1", "epsi2", "A1", "target_acceptance"):
        assert float(got[k]) == m[k], k
    assert [int(v) for v in t[t.index("Nt_learn") + 1:t.index("periods_learn")]] == m["Nt_learn"]
    assert [int(v) for v in t[t.index("periods_learn") + 1:]] == m["periods_learn"]
    p.write_text("!MALA:\n\tc0=10\n")
    with pytest.raises(ValueError):
        fmt.read_cfg(str(p))
    assert subprocess.run([exe, str(p)], stdout=subprocess.PIPE, text=True).stdout.strip() == "rc 2"


def test_ms_global_model_parse_stage(pkg, tmp_path):
    """read_ms_global_model follows read_MCMC_file_MS_Global (io_ms_global.cpp:27-360) into the fields of MCMC_files, on the
    reference's own ajAlm test input (tests/golden/reference_ms_global_model.json, make_golden_ms_global_model.py)."""
    fmt = pkg.formats
    gold = json.load(open(os.path.join(HERE, "golden", "reference_ms_global_model.json")))
    p = tmp_path / "star.model"
    p.write_text(gold["text"])
    m = fmt.read_ms_global_model(str(p))
    assert m["ID"] == "003427720" and m["Dnu"] == 119.557 and m["C_l"] == 55.0616 and m["freq_range"] == (1434.684, 3706.267)
    assert m["numax"] == -9999.0 and len(m["els"]) == 33 and [int((m["els"] == l).sum()) for l in (0, 1, 2)] == [11, 11, 11]
    assert m["freqs_ref"][0] == 1969.8199 and m["freqs_ref"][-1] == 3157.25 and all(m["relax_H"]) and all(m["relax_gamma"])
    assert m["hyper_priors"].shape == (5, 1)                      # the "extra parameters" column (the label line is skipped unread)
    assert m["eigen_params"].shape == (33, 6) and np.array_equal(m["eigen_params"][:, 0], m["els"])
    assert np.allclose(m["eigen_params"][:, 1], m["freqs_ref"], rtol=0, atol=2e-4)
    assert m["noise_params"].tolist() == [0, 0, 1, 5.446283e-31, 420.20987, 4, 24.214348, 9.9205704, 2, 1.2831577]
    assert np.array_equal(m["noise_s2"][:, 0], m["noise_params"]) and np.isinf(m["noise_s2"][3, 2])
    assert len(m["common_names"]) == gold["n_common"] == 22 and m["common_names_priors"][0] == "model_MS_Global_ajAlm_HarveyLike"
    k = m["common_names"].index("Visibility_l2")
    assert m["common_names_priors"][k] == "Gaussian" and m["modes_common"][k].tolist() == [0.53, 0.53, 0.03, -9999, -9999]
    assert m["modes_common"][m["common_names"].index("trunc_c"), 0] == 30.0
    # a short file: numax line, missing relax flags, fewer noise rows (right-aligned, -1 fill), a second '*' line is an error
    short = "#KIC =1\n!n 100.5 2.5\n! 10.0\n!! 1.5\n* 50 150\n# type\np 0 90.0\ng 1 95.5 0\n# hyper priors\n# extra\n# eigen\n 0 90 89 91 0.1 1\n# noise\n 1 2 3\n 0.5\n# s2\n 1 0 0\n# common\n trunc_c Fix 20\n"
    p.write_text(short)
    m = fmt.read_ms_global_model(str(p))
    assert m["numax"] == 100.5 and m["err_numax"] == 2.5 and m["param_type"] == ["p", "g"] and m["relax_freq"] == [True, False] and m["relax_H"] == [True, True]
    assert m["noise_params"].tolist() == [-1] * 6 + [1, 2, 3, 0.5] and m["noise_s2"][9].tolist() == [1, 0, 0] and m["noise_s2"][8, 0] == -1
    assert m["hyper_priors"].shape[0] == 0 and m["common_names"] == ["trunc_c"]
    p.write_text(short.replace("# type", "* 1 2\n# type"))
    with pytest.raises(ValueError):
        fmt.read_ms_global_model(str(p))


def test_rgb_model_file_goes_through_the_same_reader(pkg, tmp_path):
    """read_MCMC_file_asymptotic is read_MCMC_file_MS_Global (io_asymptotic.cpp:27-29): the reference's red-giant fixture
    10722175_nobias.model (tests/golden/reference_rgb_model.json) -- numax line, 16 hyper-prior rows `value prior params...`
    for the l=1 frequencies, the l=1 placeholder row of the eigen table -- and the noise values agree with the parameter vector
    the kernel-parity vectors of the same fixture were built with (reference_rgb_vectors.npz)."""
    fmt = pkg.formats
    gold = json.load(open(os.path.join(HERE, "golden", "reference_rgb_model.json")))
    p = tmp_path / "rgb.model"
    p.write_text(gold["text"])
    r = fmt.read_ms_global_model(str(p))
    assert r["ID"] == "010722175" and r["numax"] == 113.784460254 and r["err_numax"] == 0.220633701471 and r["Dnu"] == 9.54
    assert r["freq_range"] == (80.0, 128.0) and r["els"].tolist() == [0] * 5 + [1] + [2] * 5 + [3] * 2
    assert r["hyper_priors"].shape == (16, 4) and r["hyper_priors"][0].tolist() == [92.2, 0, 0, 0.1] and r["hyper_priors_names"] == ["Fix"] * 16
    assert np.all(np.diff(r["hyper_priors"][:, 0]) > 0)
    assert r["eigen_params"].shape == (12, 6) and r["eigen_params"][5].tolist() == [1, 100.0, -1, -1, -1, -1]
    assert r["common_names"] == gold["common_names"] and len(r["common_names"]) == gold["n_common"]
    g = np.load(os.path.join(HERE, "golden", "reference_rgb_vectors.npz"))
    pl, params = g["plength0"], g["params0"]
    o, nn = int(pl[:8].sum()), int(pl[8])
    assert nn == 10 and np.array_equal(r["noise_params"], params[o:o + nn])


def test_tabulated_prior_tables(pkg, tmp_path):
    """The `.priors` tables the reference ships beside its ajAlm .model (config.cpp:196-260): a 1-D PDF of a1 and a 2-D PDF of
    (a1, inclination).  Texts in tests/golden/reference_tabulated_priors.json (make_golden_ms_global_model.py)."""
    fmt = pkg.formats
    gold = json.load(open(os.path.join(HERE, "golden", "reference_tabulated_priors.json")))
    for k in ("0", "1"):
        (tmp_path / (k + ".priors")).write_text(gold[k])
    t0, t1 = fmt.read_tabulated_prior(str(tmp_path / "0.priors")), fmt.read_tabulated_prior(str(tmp_path / "1.priors"))
    assert t0["ndim"] == 1 and t0["labels"] == ["a1", "PDF"] and t0["units"] == ["(microHz)", "(no_unit)"]
    assert t0["x"].tolist() == [0, .1, .2, .3, .4, .5, .6, .7] and t0["pdf"].tolist() == [.001, .1, .2, .4, .2, .1, .05, 0]
    assert t1["ndim"] == 2 and t1["labels"] == ["a1", "Inclination", "PDF"] and t1["y"].tolist() == [0, 20, 40, 50, 60, 70, 80, 90]
    assert np.allclose(t1["x"], np.arange(10) / 10) and t1["pdf"].shape == (8, 10) and t1["pdf"][3, 2] == 0.511 and t1["pdf"][7, 4] == 0.03
    (tmp_path / "bad.priors").write_text("! a b c\nNA 1 2\n10 0.1\n")
    with pytest.raises(ValueError):
        fmt.read_tabulated_prior(str(tmp_path / "bad.priors"))


def test_tabulated_prior_matches_reference():
    """priors.hpp:logP_tabulated against the reference's own logP_tabulated (stats_dictionary.cpp:252-291) on its shipped table and
    an irregular one: in range, on the nodes, out of range (numeric_limits<double>::lowest()), negative interpolated values
    (log 0 = -inf) and the normalise flag as the reference treats it (tests/golden/reference_tabulated_logp.json)."""
    cases = json.load(open(os.path.join(HERE, "golden", "reference_tabulated_logp.json")))
    lines = ["T %d %s %s %r %d" % (len(c["tab_x"]), " ".join(repr(v) for v in c["tab_x"]), " ".join(repr(v) for v in c["tab_y"]), c["x"], c["normalise"])
             for c in cases]
    r = subprocess.run([_build()], input="\n".join(lines) + "\n", stdout=subprocess.PIPE, text=True, check=True)
    got = [float(t) for t in r.stdout.split()]
    want = [float(c["value"]) for c in cases]
    assert len(got) == len(want) > 200
    bad = [(c["x"], c["normalise"], a, b) for c, a, b in zip(cases, got, want) if not (a == b or _same(a, b))]
    assert not bad, bad[:3]
    lowest = -1.7976931348623157e308
    assert sum(1 for v in want if v == lowest) > 8 and sum(1 for v in want if np.isfinite(v) and v != lowest) > 80 and any(np.isinf(v) for v in want)


def test_remaining_binary_outputs_round_trip(pkg, tmp_path):
    """Statistical criteria, parallel-tempering log, proposal law and models of a run (Outputs::write_bin_stat_criteria
    outputs.cpp:1472-1550, write_bin_parallel_temp_params :1336-1404, write_bin_prop_params :1029-1229, write_bin_models
    :1406-1470): record sizes and orders as the reference writes them, headers with its keys, an append continues the streams."""
    F = pkg.formats
    rng = np.random.default_rng(8)
    N, nch, nv, nd = 7, 3, 4, 11
    names = ["Height_l0", "Frequency_l", "Width_l0", "Inclination"]
    # ---- statistical criteria: per sample Nchains logL, Nchains logPrior, Nchains logPosterior ----
    L, P = rng.normal(-1e5, 10, (N, nch)), rng.normal(-20, 1, (N, nch))
    stem = str(tmp_path / "star_stat_criteria")
    F.write_stat_criteria(stem, L[:4], P[:4], (L + P)[:4], Nsamples_done=4)
    F.write_stat_criteria(stem, L[4:], P[4:], (L + P)[4:], append=True)
    assert os.path.getsize(stem + ".bin") == N * 3 * nch * 8
    raw = np.fromfile(stem + ".bin", dtype="<f8")
    assert np.array_equal(raw[:nch], L[0]) and np.array_equal(raw[nch:2 * nch], P[0]) and np.array_equal(raw[3 * nch:4 * nch], L[1])
    s = F.read_stat_criteria(stem)
    assert s["Nchains"] == nch and s["Nsamples_done"] == 4
    assert np.array_equal(s["logLikelihood"], L) and np.array_equal(s["logPrior"], P) and np.array_equal(s["logPosterior"], L + P)
    hdr = open(stem + ".hdr").read().splitlines()
    assert hdr[0] == "# This is the header of the BINARY output file for the statistical information."
    assert hdr[4].startswith("! labels= logLikelihood[0]   logLikelihood[1]   logLikelihood[2]   logPrior[0]   ") and hdr[4].rstrip().endswith("logPosteriors[2]")
    # ---- parallel tempering: 14-byte records bool, int32, double, bool ----
    T = pkg.synth.tcoefs(nch, 1.7)
    att = rng.random(N) < 0.5; c0 = rng.integers(0, nch - 1, N).astype(np.int32); ps = rng.random(N); sw = att & (rng.random(N) < 0.5)
    stem = str(tmp_path / "star_parallel_tempering")
    F.write_parallel_tempering(stem, T, att, c0, ps, sw)
    assert F.PT_RECORD.itemsize == 14 and os.path.getsize(stem + ".bin") == 14 * N
    h, rec = F.read_parallel_tempering(stem)
    assert np.allclose(h["Tcoefs"], T, rtol=1e-5) and h["Nsamples_done"] == N          # (the header prints 6 significant digits, like Eigen)
    assert np.array_equal(rec["attempt_mixing"], att) and np.array_equal(rec["chain0"], c0) and np.array_equal(rec["Pswitch"], ps) and np.array_equal(rec["switched"], sw)
    b = open(stem + ".bin", "rb").read()
    assert b[0] == int(att[0]) and np.frombuffer(b[1:5], dtype="<i4")[0] == c0[0] and np.frombuffer(b[5:13], dtype="<f8")[0] == ps[0] and b[13] == int(sw[0])
    assert open(stem + ".hdr").read().splitlines()[-1] == "! labels= attempt_mixing    chain0    Pswitch    switched "
    # ---- proposal law ----
    sig = rng.random((N, nch)); mu = rng.normal(size=(N, nch, nv)); A = rng.normal(size=(N, nch, nv, nv)); cov = A @ np.swapaxes(A, 2, 3)
    pm = rng.random((N, nch)); mvd = rng.random((N, nch)) < pm
    stem = str(tmp_path / "star_proposals")
    F.write_proposals(stem, sig, mu, cov, pm, mvd, names)
    assert os.path.getsize(stem + "_moves.bin") == N * nch * 9 and os.path.getsize(stem + "_covarmats_chain-2.bin") == N * nv * nv * 8
    r = F.read_proposals(stem)
    assert r["variable_names"] == names and r["Nvars"] == nv and r["Nchains"] == nch
    for k, ref in (("sigmas", sig), ("mus", mu), ("covarmats", cov), ("Pmoves", pm), ("moveds", mvd)):
        assert np.array_equal(r[k], ref), k
    assert np.array_equal(np.fromfile(stem + "_covarmats_chain-1.bin", dtype="<f8")[:nv], cov[0, 1, 0, :])      # row by row
    assert "! Nvars= %d" % nv in open(stem + "_covarmats.hdr").read() and "Pmove[0:Nchains-1]" in open(stem + "_moves.hdr").read()
    # ---- models ----
    M = rng.random((N, nch, nd))
    stem = str(tmp_path / "star")
    F.write_models(stem, M[:3]); F.write_models(stem, M[3:], append=True)
    assert np.array_equal(F.read_models(stem, nch, nd), M)
