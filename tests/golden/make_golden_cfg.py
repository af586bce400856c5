"""Golden values of the reference's shipped control file Config/default/config_default.cfg as tamcmc-c_b200/formats.py reads it
(read_cfg / mala_config).  Only the parsed values are stored.  Run in the build container (needs /root/reference):
    python tests/golden/make_golden_cfg.py"""
import importlib.util
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("formats", os.path.join(HERE, "..", "..", "tamcmc-c_b200", "formats.py"))
fmt = importlib.util.module_from_spec(spec)
spec.loader.exec_module(fmt)
g = fmt.read_cfg("/root/reference/Config/default/config_default.cfg")
mala = fmt.mala_config(g)
# the values mcmc_driver.hpp cites from this file (config_default.cfg:11-29)
assert mala["target_acceptance"] == 0.234 and mala["c0"] == 10 and mala["epsilon1"] == 1e-12 and mala["A1"] == 1e14
assert mala["Nt_learn"] == [1000, 1500, 100000] and mala["periods_learn"] == [1, 1] and mala["Nchains"] == 5 and mala["lambda_temp"] == 3.5
out = {"source": "Config/default/config_default.cfg", "groups": sorted(g), "MALA": mala,
       "Modeling": {k: g["Modeling"][k] for k in ("prior_fct_name", "model_fct_name", "likelihood_fct_name", "likelihood_params")},
       "Data": {k: g["Data"][k] for k in ("x_col", "y_col", "ysig_col")}}
json.dump(out, open(os.path.join(HERE, "reference_cfg_default.json"), "w"), indent=1)
print(out)
