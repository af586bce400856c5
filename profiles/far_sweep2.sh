run() { tag=$1; shift; env "$@" python bench.py --no-cpu-baseline --steps 500 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$tag', round(d['value']), 'evals/s', round(d['ms_per_step']*1e3,1),'us/step kernel',round(r['kernel_ms']*1e3,1),'expand',round(r['expand_kernel_ms']*1e3,1),'frac',round(r['frac'],3),'e2e',round(d['e2e']['value']))"; }
run stag300 TAMCMC_GPU_STAGGER_NS=300
run stag1000 TAMCMC_GPU_STAGGER_NS=1000
run stag2000 TAMCMC_GPU_STAGGER_NS=2000
run pdl TAMCMC_GPU_PDL=1
run tile768 TAMCMC_GPU_TILE=768
