"""A/B runner for library switches (run on the GPU box): repeats `bench.py --no-cpu-baseline` under each environment
variant and prints step time, e2e step time and the event-timed kernel durations side by side.
  python profiles/ab.py [--steps N] [--reps R] [--stars S] name1:VAR=val,VAR2=val name2: ...
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
steps, reps, stars = 400, 2, 1
variants = []
while args:
    a = args.pop(0)
    if a == "--steps": steps = int(args.pop(0))
    elif a == "--reps": reps = int(args.pop(0))
    elif a == "--stars": stars = int(args.pop(0))
    else:
        name, _, envs = a.partition(":")
        variants.append((name, dict(e.split("=", 1) for e in envs.split(",") if e)))
for r in range(reps):
    for name, env in variants:
        e = dict(os.environ); e.update(env)
        p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(steps), "--warmup", "10", "--no-cpu-baseline",
                            "--stars-per-gpu", str(stars)], env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        if p.returncode != 0:
            print(json.dumps({"variant": name, "error": p.stderr[-400:]})); continue
        j = json.loads(p.stdout.strip().splitlines()[-1])
        print(json.dumps({"variant": name, "rep": r, "us_per_step": round(1e3 * j["ms_per_step"], 2), "evals_per_s": round(j["value"]),
                          "e2e_us": round(1e3 * j["e2e"]["ms_per_step"], 2), "fused_us": round(1e3 * j["roofline"]["kernel_ms"], 2),
                          "expand_us": round(1e3 * j["roofline"]["expand_kernel_ms"], 2), "frac": round(j["roofline"]["frac"], 3),
                          "sm_mhz": j["clocks"]["sm_mhz"],
                          "c3_us": round(1e3 * j["extra"]["c3_bin_sharded"]["ms_per_step"], 2) if "c3_bin_sharded" in j.get("extra", {}) and "ms_per_step" in j["extra"]["c3_bin_sharded"] else j.get("extra", {}).get("c3_bin_sharded"),
                          "c3_err": j.get("extra", {}).get("c3_bin_sharded", {}).get("max_rel_err_vs_reference_logL"),
                          "c5_evals": round(j["extra"]["c5_star_sharded"]["value"]) if "value" in j.get("extra", {}).get("c5_star_sharded", {}) else j.get("extra", {}).get("c5_star_sharded")}), flush=True)
