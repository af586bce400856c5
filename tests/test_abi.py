"""The C-ABI shared library loads and exports every symbol include/tamcmc_gpu.h declares.
No compute calls here (no GPU in the build container); on a machine without a CUDA device the
library must FAIL LOUDLY -- there is no CPU fallback behind the ABI."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "tamcmc_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tamcmc_(?:gpu|host|alm)_[a-zA-Z0-9_]+)\s*\(", src)))


def test_header_functions_match_binding_list(pkg):
    assert _declared_functions() == sorted(pkg.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(pkg):
    L = C.CDLL(pkg.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(L, name), name


def test_abi_version_and_strerror(pkg):
    L = pkg.lib()
    assert L.tamcmc_gpu_abi_version() == 2
    for rc in range(7):
        assert len(L.tamcmc_gpu_strerror(rc)) > 0


def test_struct_layout_matches_header(pkg):
    # tamcmc_gpu_star: int, int[11], int, 2 pointers, 3 longs, 3 doubles, 1 pointer (LP64)
    assert C.sizeof(pkg.StarStruct) == 4 + 44 + 4 + 4 + 8 * 2 + 8 * 3 + 8 * 3 + 8


def test_argument_validation_without_device(pkg):
    L = pkg.lib()
    h = C.c_void_p()
    assert L.tamcmc_gpu_create(0, 0, None, 1, None, 1.0, 0, C.byref(h)) == pkg.ERR_ARG
    assert L.tamcmc_gpu_eval(None, None, None, None, None) == pkg.ERR_ARG
    assert L.tamcmc_gpu_params_stride(None) == 0


def test_no_cpu_fallback(pkg):
    """Without a usable CUDA device, create() must return TAMCMC_ERR_CUDA (never evaluate on the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present; the loud-failure path is exercised in the CPU container")
    x = pkg.synth.freq_axis(4096, 900.0, 0.1)
    params, pl = pkg.synth.classic_params(np.random.default_rng(0), Nmax=4, lmax=2, f0=950.0, dnu=80.0)
    with pytest.raises(pkg.TamcmcError) as ei:
        pkg.Context(pkg.Star(3, pl, len(params), x, np.ones_like(x)), 2, [1.0, 2.0])
    assert ei.value.status == pkg.ERR_CUDA
    with pytest.raises(pkg.TamcmcError):
        pkg.fp64_peak(0)


def test_product_never_references_the_oracle():
    """The product tree must not import, link or execute anything under oracle/."""
    bad = []
    for base in ("tamcmc-c_b200", "include"):
        for dp, dn, fn in os.walk(os.path.join(ROOT, base)):
            if "build" in dp.split(os.sep):
                continue
            for f in fn:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".c", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if "oracle/" in txt or "tamcmc_oracle" in txt or "_oracle" in txt:
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_header_is_plain_c_and_links_from_c(pkg, tmp_path):
    """The boundary is a C ABI: the header compiles as strict C99 and a C program links against the library
    (no C++ runtime, no torch types in any signature)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi_check.c"
    src.write_text('#include "tamcmc_gpu.h"\n#include <stdio.h>\n'
                   "int main(void) {\n"
                   "  tamcmc_gpu_star s; tamcmc_gpu_ctx *c = 0; (void)s;\n"
                   "  if (tamcmc_gpu_create(0, 0, 0, 1, 0, 1.0, TAMCMC_LIKELIHOOD_CHI22P, &c) != TAMCMC_ERR_ARG) return 1;\n"
                   '  printf("%d %s\\n", tamcmc_gpu_abi_version(), tamcmc_gpu_strerror(TAMCMC_ERR_WINDOW));\n'
                   "  return sizeof(s) == 128 ? 0 : 2;\n}\n")
    libdir = os.path.join(root, "tamcmc-c_b200")
    exe = tmp_path / "abi_check"
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I" + os.path.join(root, "include"), "-o", str(exe), str(src),
                           "-L" + libdir, "-ltamcmc_gpu", "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + libdir,
                           "-Wl,-rpath,/usr/local/cuda/lib64"])
    r = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout
    assert r.stdout.split()[0] == "2"
