"""Golden vector for the binary-output header (tamcmc-c_b200/formats.py: params_header_text / parse_params_header): the text of the
header the reference itself wrote for its Gaussian-envelope example (a 465-byte metadata file shipped with the reference's
tools) and the values it encodes.  Run in the build container (needs /root/reference):
    python tests/golden/make_golden_params_hdr.py"""
import json
import os

SRC = "/root/reference/tools/convert_fit2prior_table/test_data/10280410_Gaussfit/outputs/10280410_Gaussfit_A_params.hdr"
text = open(SRC).read()
vals = {}
for line in text.splitlines():
    if line.startswith("!"):
        k, _, v = line[1:].partition("=")
        vals[k.strip()] = v.split()
out = {"source": SRC.replace("/root/reference/", ""), "text": text, "tokens": vals}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_params_hdr.json"), "w"), indent=1)
print(vals)
