"""The C++ host mirror of the reference's Model_def (tamcmc-c_b200/host/model_def_gpu.hpp) compiled with g++ against
the C ABI: without a GPU it must fail loudly (no CPU fallback); on the GPU it must match the oracle."""
import os
import subprocess

import numpy as np
import pytest

import _cases

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EXE = os.path.join(HERE, "cpp", "test_model_def_gpu")
LIBDIR = os.path.join(ROOT, "tamcmc-c_b200")


def _build():
    src = os.path.join(HERE, "cpp", "test_model_def_gpu.cpp")
    hdr = os.path.join(LIBDIR, "host", "model_def_gpu.hpp")
    abi = os.path.join(ROOT, "include", "tamcmc_gpu.h")
    if os.path.exists(EXE) and os.path.getmtime(EXE) > max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(abi)):
        return EXE
    cuda_lib = "/usr/local/cuda/lib64"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-o", EXE, src, "-L" + LIBDIR, "-ltamcmc_gpu",
                           "-L" + cuda_lib, "-lcudart", "-Wl,-rpath," + LIBDIR, "-Wl,-rpath," + cuda_lib])
    return EXE


def _build_driver():
    exe = os.path.join(HERE, "cpp", "test_mcmc_driver")
    src = os.path.join(HERE, "cpp", "test_mcmc_driver.cpp")
    hdr = os.path.join(LIBDIR, "host", "mcmc_driver.hpp")
    import _oracle
    _oracle.build()
    odir = os.path.join(ROOT, "oracle", "_ref")
    abi = os.path.join(ROOT, "include", "tamcmc_gpu.h")
    if os.path.exists(exe) and os.path.getmtime(exe) > max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(abi)):
        return exe
    cuda_lib = "/usr/local/cuda/lib64"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-mavx2", "-Wall", "-o", exe, src, "-L" + LIBDIR, "-ltamcmc_gpu", "-L" + odir, "-ltamcmc_oracle",
                           "-L" + cuda_lib, "-lcudart", "-fopenmp", "-Wl,-rpath," + LIBDIR, "-Wl,-rpath," + odir, "-Wl,-rpath," + cuda_lib])
    return exe


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_cpp_mirror_builds_and_fails_loudly_without_gpu(pkg):
    pkg.lib()          # the shared library must exist (built by __graft_entry__.build())
    exe = _build()
    if _have_gpu():
        pytest.skip("CUDA device present: covered by the gpu-marked test")
    r = subprocess.run([exe, "nogpu"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    assert "status 3" in r.stdout      # TAMCMC_ERR_CUDA


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle(pkg, oracle, tmp_path):
    exe = _build()
    Nmodels = 5
    params, pl, x = _cases.ms_case(pkg.synth, 3, seed=4, N=20000, asym=12.0)
    rc, M = oracle.call_model(3, params, pl, x)
    assert rc == 0
    rng = np.random.default_rng(8)
    y = pkg.synth.chi2_2dof_spectrum(rng, M)
    P = pkg.synth.perturb_chains(rng, params, pl, Nmodels)
    T = pkg.synth.tcoefs(Nmodels, 1.7)
    rc, L = oracle.eval_chains(3, P, pl, x, y, T)
    assert rc == 0
    logPrior = np.array([-3.0, 0.0, -np.inf, -1.5, -np.inf])
    f = tmp_path / "case.bin"
    hdr = np.concatenate([[3, len(x), Nmodels, len(params), 1.0], pl.astype(float)])
    with open(f, "wb") as fh:
        for a in (hdr, x, y, T, P.ravel(), logPrior, L, M):
            fh.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    r = subprocess.run([exe, str(f)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout


def _build_rgb():
    exe = os.path.join(HERE, "cpp", "test_model_def_rgb")
    src = os.path.join(HERE, "cpp", "test_model_def_rgb.cpp")
    deps = [src, os.path.join(LIBDIR, "host", "model_def_rgb.hpp"), os.path.join(LIBDIR, "host", "model_def_gpu.hpp"), os.path.join(ROOT, "include", "tamcmc_gpu.h")]
    if os.path.exists(exe) and os.path.getmtime(exe) > max(os.path.getmtime(d) for d in deps):
        return exe
    cuda_lib = "/usr/local/cuda/lib64"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-o", exe, src, "-L" + LIBDIR, "-ltamcmc_gpu",
                           "-L" + cuda_lib, "-lcudart", "-Wl,-rpath," + LIBDIR, "-Wl,-rpath," + cuda_lib])
    return exe


def test_cpp_rgb_mirror_builds_and_fails_loudly_without_gpu(pkg):
    pkg.lib()
    exe = _build_rgb()
    if _have_gpu():
        pytest.skip("CUDA device present: covered by the gpu-marked test")
    r = subprocess.run([exe, "nogpu"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "status 3" in r.stdout, r.stdout


@pytest.mark.gpu
def test_cpp_rgb_mirror_matches_reference(pkg, oracle, tmp_path):
    """ModelDefRGB (host/model_def_rgb.hpp): reference parameter vectors of the red-giant fixture -> device set-up + evaluation, against
    the oracle's log-likelihoods of the host expander's rows and the spectrum the REFERENCE's model function returned."""
    exe = _build_rgb()
    G = np.load(os.path.join(HERE, "golden", "reference_rgb_vectors.npz"))
    x, y = G["x"], G["y"]
    pl = G["plength0"]
    cases = [0, 1, 2, 3, 0]
    assert all(np.array_equal(pl, G["plength%d" % i]) for i in cases)
    P = np.stack([G["params%d" % i] for i in cases])
    P[4, :int(pl[0])] *= 1.03
    Nmodels, cap, nn = len(cases), 120, int(pl[8])
    T = pkg.synth.tcoefs(Nmodels, 3.5)
    rows = np.stack([pkg.expand_rgb_v4(25, P[m], pl, x[2] - x[1], cap)[0] for m in range(Nmodels)])
    rc, L = oracle.mode_table_eval_chains(rows, nn, 1, x, y, T)
    assert rc == 0
    logPrior = np.array([-2.0, -np.inf, 0.0, -1.0, -0.5])
    f = tmp_path / "case.bin"
    hdr = np.concatenate([[25, len(x), Nmodels, P.shape[1], cap], pl.astype(float)])
    with open(f, "wb") as fh:
        for a in (hdr, x, y, T, P.ravel(), logPrior, L, G["model0"]):
            fh.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    r = subprocess.run([exe, str(f)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    print(r.stdout)
    assert r.returncode == 0 and "model_def_rgb: ok" in r.stdout, r.stdout


def _build_rgb_driver():
    exe = os.path.join(HERE, "cpp", "test_rgb_driver")
    src = os.path.join(HERE, "cpp", "test_rgb_driver.cpp")
    deps = [src, os.path.join(LIBDIR, "host", "mcmc_driver.hpp"), os.path.join(LIBDIR, "host", "model_def_gpu.hpp"), os.path.join(ROOT, "include", "tamcmc_gpu.h")]
    if os.path.exists(exe) and os.path.getmtime(exe) > max(os.path.getmtime(d) for d in deps):
        return exe
    cuda_lib = "/usr/local/cuda/lib64"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-fopenmp", "-o", exe, src, "-L" + LIBDIR, "-ltamcmc_gpu",
                           "-L" + cuda_lib, "-lcudart", "-Wl,-rpath," + LIBDIR, "-Wl,-rpath," + cuda_lib])
    return exe


def rgb_driver_case(path, nchains=5):
    """the red-giant fixture as a fit: heights, l=0 frequencies, the four mixed-mode hyper-parameters, the two rotation rates and the
    inclination relaxed inside uniform boxes"""
    G = np.load(os.path.join(HERE, "golden", "reference_rgb_vectors.npz"))
    x, y, params, pl = G["x"], G["y"], G["params0"], G["plength0"]
    Nmax, lmax, Nfl0, Nfl1, Nfl2, Nfl3, Nsplit, Nwidth, Nnoise = [int(v) for v in pl[:9]]
    o0 = Nmax + lmax
    o1 = o0 + Nfl0
    o_s = o0 + Nfl0 + Nfl1 + Nfl2 + Nfl3
    o_inc = o_s + Nsplit + Nwidth + Nnoise
    relax, err, lo, hi = [], [], [], []
    for k in range(Nmax):                                  # heights
        relax.append(k); err.append(0.05 * abs(params[k])); lo.append(0.2 * abs(params[k])); hi.append(5 * abs(params[k]))
    for k in range(Nfl0):                                  # l=0 frequencies
        relax.append(o0 + k); err.append(0.01); lo.append(params[o0 + k] - 0.5); hi.append(params[o0 + k] + 0.5)
    for k, (e, a, b) in enumerate([(0.005, -0.3, 0.3), (0.02, params[o1 + 1] - 2, params[o1 + 1] + 2), (0.01, 0.0, 1.0), (0.005, 0.05, 0.6)]):
        relax.append(o1 + k); err.append(e); lo.append(a); hi.append(b)        # delta0l, DPl, alpha_g, q
    relax += [o_s, o_s + 1]; err += [0.01, 0.01]; lo += [0.0, 0.0]; hi += [1.0, 2.0]        # rot_env, rot_core
    relax.append(o_inc); err.append(1.0); lo.append(0.0); hi.append(90.0)
    hdr = np.concatenate([[25, len(x), nchains, len(params), 140, len(relax)], pl.astype(float)])
    with open(path, "wb") as fh:
        for a in (hdr, x, y, params, relax, err, lo, hi):
            fh.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return path


def test_cpp_rgb_driver_builds(pkg):
    pkg.lib()
    _build_rgb_driver()


@pytest.mark.gpu
def test_cpp_driver_samples_a_red_giant_fit_with_the_device_setup(pkg, tmp_path):
    """BASELINE C1 (the reference's RGBtests preset on fixture 10722175, 5 chains): the C++ driver with one tamcmc_gpu_rgb_expand +
    one tamcmc_gpu_eval per step.  Same seed -> identical chains; the chain with the host solver is the same chain."""
    exe = _build_rgb_driver()
    f = rgb_driver_case(str(tmp_path / "case.bin"))
    r = subprocess.run([exe, f, "600", "host"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    print(r.stdout)
    assert r.returncode == 0 and "rgb driver: ok" in r.stdout, r.stdout


def test_cpp_driver_builds(pkg):
    pkg.lib()
    _build_driver()


@pytest.mark.gpu
def test_cpp_driver_posterior_consistent_with_cpu_oracle(pkg, oracle, tmp_path):
    """Fixed-seed adaptive Metropolis + parallel tempering (host/mcmc_driver.hpp) with the GPU likelihood vs the same
    driver with the CPU oracle likelihood: same posterior summaries."""
    exe = _build_driver()
    Nmodels = 4
    params, pl, x = _cases.ms_case(pkg.synth, 3, seed=6, N=6000, Nmax=3, lmax=2, trunc_c=10.0)
    rc, M = oracle.call_model(3, params, pl, x)
    assert rc == 0
    rng = np.random.default_rng(12)
    y = pkg.synth.chi2_2dof_spectrum(rng, M)
    P = np.tile(params, (Nmodels, 1))
    T = pkg.synth.tcoefs(Nmodels, 1.7)
    f = tmp_path / "case.bin"
    hdr = np.concatenate([[3, len(x), Nmodels, len(params), 1.0], pl.astype(float)])
    with open(f, "wb") as fh:
        for a in (hdr, x, y, T, P.ravel()):
            fh.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    r = subprocess.run([exe, str(f), "3000"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    print(r.stdout)
    assert r.returncode == 0, r.stdout


@pytest.mark.gpu
def test_cpp_batch_driver_reproduces_single_star_chain(pkg, oracle, tmp_path):
    """BatchDriver (many independent stars, one batched tamcmc_gpu_eval per step): star 0 of a 5-star batch reproduces the
    single-star run bit for bit (same seed, same data), whatever the OpenMP thread count."""
    exe = _build_driver()
    Nmodels = 3
    params, pl, x = _cases.ms_case(pkg.synth, 3, seed=8, N=5000, Nmax=3, lmax=2, trunc_c=10.0)
    rc, M = oracle.call_model(3, params, pl, x)
    assert rc == 0
    y = pkg.synth.chi2_2dof_spectrum(np.random.default_rng(3), M)
    T = pkg.synth.tcoefs(Nmodels, 1.7)
    f = tmp_path / "case.bin"
    hdr = np.concatenate([[3, len(x), Nmodels, len(params), 1.0], pl.astype(float)])
    with open(f, "wb") as fh:
        for a in (hdr, x, y, T, np.tile(params, (Nmodels, 1)).ravel()):
            fh.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    r = subprocess.run([exe, str(f), "800", "batch", "5"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    print(r.stdout)
    assert r.returncode == 0 and '"star0_matches_single_run": true' in r.stdout, r.stdout


def test_cpp_driver_restores_a_previous_run(tmp_path):
    """Driver::restore_proposal / restore_variables (MALA.cpp:191-246, do_restore): CPU-only, analytic Gaussian likelihood."""
    exe, src = os.path.join(HERE, "cpp", "test_driver_restore"), os.path.join(HERE, "cpp", "test_driver_restore.cpp")
    hdrs = [os.path.join(HERE, "..", "tamcmc-c_b200", "host", h) for h in ("mcmc_driver.hpp", "outputs.hpp")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in hdrs]):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-fopenmp", "-o", exe, src])
    r = subprocess.run([exe, str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "identical state" in r.stdout and "restart through the restore files: ok" in r.stdout, r.stdout
    assert sorted(os.listdir(tmp_path)) == ["star_restore_A_%d.dat" % n for n in (1, 2, 3)]


def test_cpp_driver_samples_the_right_law():
    """Adaptive Metropolis + parallel tempering (MALA.cpp:296-319, 397-461, 463-553) on a correlated Gaussian with known
    covariance: chain 0 mean/covariance, chain m covariance = T_m x that, acceptance at the Robbins-Monro target, PT swap rate
    against its expectation; and a failing evaluator rejects every proposal and reports its status (CPU only)."""
    exe, src = os.path.join(HERE, "cpp", "test_driver_law"), os.path.join(HERE, "cpp", "test_driver_law.cpp")
    hdr = os.path.join(HERE, "..", "tamcmc-c_b200", "host", "mcmc_driver.hpp")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-fopenmp", "-o", exe, src])
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    print(r.stdout)
    assert r.returncode == 0 and "driver law: ok" in r.stdout, r.stdout
