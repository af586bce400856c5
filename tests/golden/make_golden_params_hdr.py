"""Golden vectors for the on-disk outputs (tamcmc-c_b200/formats.py, host/outputs.hpp): the small ASCII files the reference
itself wrote for its Gaussian-envelope example (shipped with the reference's tools as test data) -- the binary-output header,
the acceptance log and the three restore files.  Run in the build container (needs /root/reference):
    python tests/golden/make_golden_params_hdr.py"""
import json
import os

ROOT = "/root/reference/tools/convert_fit2prior_table/test_data/10280410_Gaussfit/"
HERE = os.path.dirname(os.path.abspath(__file__))
SRC = ROOT + "outputs/10280410_Gaussfit_A_params.hdr"
text = open(SRC).read()
vals = {}
for line in text.splitlines():
    if line.startswith("!"):
        k, _, v = line[1:].partition("=")
        vals[k.strip()] = v.split()
out = {"source": SRC.replace("/root/reference/", ""), "text": text, "tokens": vals}
json.dump(out, open(os.path.join(HERE, "reference_params_hdr.json"), "w"), indent=1)
more = {"acceptance": open(ROOT + "outputs/10280410_Gaussfit_A_acceptance.txt").read(),
        "restore": {str(n): open(ROOT + "restore/10280410_Gaussfit_restore_A_%d.dat" % n).read() for n in (1, 2, 3)}}
json.dump(more, open(os.path.join(HERE, "reference_outputs_10280410.json"), "w"), indent=1)
print(vals, {k: len(v) for k, v in more["restore"].items()})
