#!/usr/bin/env python
"""Writes the committed SASS evidence for every kernel of libtamcmc_gpu.so (no GPU needed: cuobjdump on the objects):

  profiles/<tag>/sass/<kernel>.sass        full listing (address, instruction) of each kernel
  profiles/<tag>/sass/SUMMARY.txt          per-kernel instruction counts by mnemonic, registers / shared memory from
                                           ptxas -v, and the Blackwell/Hopper-era mnemonics that prove what is used:
                                           UBLKCP (TMA bulk copy cp.async.bulk), SYNCS (mbarrier), DFMA/DMUL (FP64 pipe)

usage: python profiles/make_sass_listing.py r1"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tamcmc-c_b200", "csrc", "build")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    out = os.path.join(ROOT, "profiles", tag, "sass")
    os.makedirs(out, exist_ok=True)
    summary = []
    for obj in ("expand.o", "whittle.o", "whittle_tiles.o", "rgb_device.o"):
        txt = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], stdout=subprocess.PIPE, text=True, check=True).stdout
        cur, body = None, collections.OrderedDict()
        for line in txt.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1)
                body[cur] = []
                continue
            m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?)\s*;?\s*/\*", line)
            if m and cur:
                body[cur].append((m.group(1), m.group(2).rstrip(" ;")))
        for fn, ins in body.items():
            short = re.sub(r"^_ZN?\d*_GLOBAL__N__[0-9a-f]+_\d+_\w+_cu_[0-9a-f]+", "", fn)
            short = re.sub(r"^_Z?\d*", "", short)
            name = re.sub(r"[^A-Za-z0-9_]", "_", short)[:60]
            if "whittle_kernelILb1" in fn:
                name = "tamcmc_whittle_kernel_write_model"
            elif "whittle_kernelILb0" in fn:
                name = "tamcmc_whittle_kernel"
            elif "expand_kernel" in fn:
                name = "tamcmc_expand_kernel"
            elif "dfma" in fn:
                name = "tamcmc_dfma_kernel"
            elif "lnx" in fn:
                name = "tamcmc_lnx_kernel"
            with open(os.path.join(out, name + ".sass"), "w") as f:
                f.write("// %s  (%s, sm_100a, cuobjdump -sass)\n" % (fn, obj))
                for a, i in ins:
                    f.write("/*%s*/ %s\n" % (a, i))
            cnt = collections.Counter()
            for _, i in ins:
                toks = i.split()
                op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
                cnt[op.split(".")[0]] += 1
            key = ["DFMA", "DMUL", "DADD", "MUFU", "UBLKCP", "SYNCS", "LDS", "LDG", "STG", "ATOMS", "ATOMG", "RED", "SHFL", "BAR", "LDL", "STL",
                   "HMMA", "UTCHMMA", "UTMALDG"]
            summary.append("%-36s %6d instructions   " % (name, len(ins)) + "  ".join("%s=%d" % (k, cnt[k]) for k in key if cnt[k]))
    for log in ("expand.ptxas.log", "whittle.ptxas.log"):
        p = os.path.join(BUILD, log)
        if os.path.exists(p):
            for line in open(p):
                if "Used" in line or "Compiling entry" in line or "stack frame" in line:
                    summary.append("ptxas %s: %s" % (log.split(".")[0], line.strip()))
    with open(os.path.join(out, "SUMMARY.txt"), "w") as f:
        f.write("SASS evidence (sm_100a).  UBLKCP = cp.async.bulk (TMA bulk copy) issued by the producer warp; SYNCS = mbarrier\n"
                "arrive/try_wait; DFMA/DMUL/DADD = FP64 pipe.  No HMMA/UTC*MMA: the path is not a dense contraction (DESIGN.md 3).\n\n")
        f.write("\n".join(summary) + "\n")
    print("\n".join(summary))


if __name__ == "__main__":
    main()
