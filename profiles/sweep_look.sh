for v in "L3E3:" "L3E2:TAMCMC_GPU_LOOK_END=2" "L3E1:TAMCMC_GPU_LOOK_END=1" "L2E2:TAMCMC_GPU_LOOK=2,TAMCMC_GPU_LOOK_END=2" "L2E1:TAMCMC_GPU_LOOK=2,TAMCMC_GPU_LOOK_END=1" "T768:TAMCMC_GPU_TILE=768" "PDL:TAMCMC_GPU_PDL=1" "L3E3b:"; do
  name=${v%%:*}; envs=${v#*:}
  out=$(env $(echo $envs | tr ',' ' ') timeout 120 python bench.py --steps 400 --warmup 10 --no-cpu-baseline --no-extra 2>/dev/null | tail -1)
  echo "$name $(echo $out | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(round(1e3*j['ms_per_step'],2), round(j['value']), 'fused', round(1e3*j['roofline']['kernel_ms'],2), 'expand', round(1e3*j['roofline']['expand_kernel_ms'],2))")"
done
