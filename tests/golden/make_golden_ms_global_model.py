"""Golden vector for formats.read_ms_global_model: the reference's own ajAlm test input (a 118-line .model data file) and the
facts a reader must get out of it, asserted here against the file.  Run in the build container (needs /root/reference):
    python tests/golden/make_golden_ms_global_model.py"""
import importlib.util
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/test/inputs/kplr003427720_kasoc-psd_slc_v1_ajAlm_gate.model"
spec = importlib.util.spec_from_file_location("formats", os.path.join(HERE, "..", "..", "tamcmc-c_b200", "formats.py"))
fmt = importlib.util.module_from_spec(spec)
spec.loader.exec_module(fmt)
m = fmt.read_ms_global_model(SRC)
text = open(SRC).read()
# facts read off the file by eye ("33 modes l<=2", mode list at .model:6-38)
assert m["ID"] == "003427720" and m["Dnu"] == 119.557 and m["C_l"] == 55.0616 and m["freq_range"] == (1434.684, 3706.267)
assert len(m["els"]) == 33 and [int((m["els"] == l).sum()) for l in (0, 1, 2)] == [11, 11, 11]
assert m["freqs_ref"][0] == 1969.8199 and m["freqs_ref"][-1] == 3157.25 and all(m["relax_freq"]) and set(m["param_type"]) == {"p"}
assert m["hyper_priors"].shape == (5, 1) and not m["hyper_priors"].any()
assert m["eigen_params"].shape == (33, 6) and m["eigen_params"][0].tolist() == [0, 1969.81995, 1965.20764, 1972.13269, 1.24707, 0.36161]
assert m["noise_params"].tolist() == [0, 0, 1, 5.446283e-31, 420.20987, 4, 24.214348, 9.9205704, 2, 1.2831577]
assert m["noise_s2"].shape == (10, 3) and m["noise_s2"][3, 2] == float("inf") and m["noise_s2"][9].tolist() == [1.2831577, 0.0059365905, 0.0059641842]
assert m["common_names"][0] == "model_fullname" and m["common_names_priors"][0] == "model_MS_Global_ajAlm_HarveyLike"
assert m["common_names"][-1] == "trunc_c" and m["modes_common"][-1].tolist() == [30.0, -9999, -9999, -9999, -9999]
k = m["common_names"].index("epsilon_0")
assert m["common_names_priors"][k] == "Jeffreys" and m["modes_common"][k].tolist() == [0.005, 0.001, 0.01, -9999, -9999]
json.dump({"source": SRC.replace("/root/reference/", ""), "text": text, "n_common": len(m["common_names"])},
          open(os.path.join(HERE, "reference_ms_global_model.json"), "w"), indent=1)
print("ok", len(m["common_names"]), m["common_names"])

# ---- the red-giant dialect: same reader (read_MCMC_file_asymptotic -> read_MCMC_file_MS_Global), the reference's fixture 10722175
SRC2 = "/root/reference/test/inputs/RGB/v1.86.0/10722175_nobias.model"
r = fmt.read_ms_global_model(SRC2)
assert r["ID"] == "010722175" and r["numax"] == 113.784460254 and r["err_numax"] == 0.220633701471 and r["Dnu"] == 9.54 and r["C_l"] == 1.2867
assert r["freq_range"] == (80.0, 128.0) and r["els"].tolist() == [0] * 5 + [1] + [2] * 5 + [3] * 2
assert r["hyper_priors"].shape == (16, 4) and r["hyper_priors"][0].tolist() == [92.2, 0, 0, 0.1] and set(r["hyper_priors_names"]) == {"Fix"}
assert r["eigen_params"].shape == (12, 6) and r["eigen_params"][5].tolist() == [1, 100.0, -1, -1, -1, -1]
json.dump({"source": SRC2.replace("/root/reference/", ""), "text": open(SRC2).read(), "n_common": len(r["common_names"]),
           "common_names": r["common_names"]}, open(os.path.join(HERE, "reference_rgb_model.json"), "w"), indent=1)
print("rgb ok", r["hyper_priors"].shape, r["eigen_params"].shape, len(r["common_names"]), r["common_names"][:6], r["noise_params"].tolist())

# ---- the two tabulated-prior tables shipped beside the ajAlm .model (tiny data files)
tabs = {}
for k in (0, 1):
    pth = "/root/reference/test/inputs/kplr003427720_kasoc-psd_slc_v1_ajAlm_gate_%d.priors" % k
    tabs[str(k)] = open(pth).read()
t0, t1 = [fmt.read_tabulated_prior("/root/reference/test/inputs/kplr003427720_kasoc-psd_slc_v1_ajAlm_gate_%d.priors" % k) for k in (0, 1)]
assert t0["ndim"] == 1 and t0["labels"] == ["a1", "PDF"] and t0["x"].tolist() == [0, .1, .2, .3, .4, .5, .6, .7] and t0["pdf"].tolist() == [.001, .1, .2, .4, .2, .1, .05, 0]
assert t1["ndim"] == 2 and t1["labels"] == ["a1", "Inclination", "PDF"] and t1["y"].tolist() == [0, 20, 40, 50, 60, 70, 80, 90]
assert t1["x"].size == 10 and t1["pdf"].shape == (8, 10) and t1["pdf"][3, 2] == 0.511 and t1["pdf"][7, 4] == 0.03
json.dump(tabs, open(os.path.join(HERE, "reference_tabulated_priors.json"), "w"), indent=1)
print("tabulated ok")
