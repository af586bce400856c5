#!/usr/bin/env python
"""Longer runs of the C++ driver on the red-giant fixture (tests/cpp/test_rgb_driver.cpp): C1 (5 chains) and C4 (10 chains), device set-up +
evaluation per step; prints the program's JSON lines.  usage: python profiles/rgb_driver_run.py [nsteps]"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_host_cpp as T  # noqa: E402

nsteps = sys.argv[1] if len(sys.argv) > 1 else "5000"
exe = T._build_rgb_driver()
d = tempfile.mkdtemp()
for nch in (5, 10):
    f = T.rgb_driver_case(os.path.join(d, "case%d.bin" % nch), nchains=nch)
    r = subprocess.run([exe, f, nsteps, "host"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    print(r.stdout.strip())
