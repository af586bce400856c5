F='import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d.get("roofline",{}); print(d.get("config",{}).get("workload","")[:12] if isinstance(d.get("config"),dict) else d.get("config"), d.get("value"), d.get("ms_per_step"), r.get("frac"), (d.get("e2e") or {}).get("value"), d.get("kernel_ms"))'
for i in 1 2; do
python bench.py --steps 400 --warmup 5 --no-cpu-baseline | python -c "$F"
TAMCMC_GPU_LIB=$PWD/tamcmc-c_b200/libtamcmc_gpu_old.so python bench.py --steps 400 --warmup 5 --no-cpu-baseline | python -c "$F"
done
python profiles/bench_configs.py --configs c1,c4,c3,c5 --steps 50 | python -c "$F"
TAMCMC_GPU_LIB=$PWD/tamcmc-c_b200/libtamcmc_gpu_old.so python profiles/bench_configs.py --configs c1,c4,c3,c5 --steps 50 | python -c "$F"
