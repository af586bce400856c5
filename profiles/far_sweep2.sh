# quick A/B harness on the bench workload (C2) and on 32 stars per launch; edit the `run` lines
run() { tag=$1; shift; env "$@" python bench.py --no-cpu-baseline --steps ${STEPS:-500} $ARGS 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$tag', round(d['value']), 'evals/s', round(d['ms_per_step']*1e3,1),'us/step kernel',round(r['kernel_ms']*1e3,1),'expand',round(r['expand_kernel_ms']*1e3,1),'frac',round(r['frac'],3),'e2e',round(d['e2e']['value']))"; }
L=$PWD/tamcmc-c_b200
for a in "" "--stars-per-gpu 32"; do
ARGS=$a; STEPS=500; [ -n "$a" ] && STEPS=60
echo "== $a"
run claim_early X=1
run claim_after_release TAMCMC_GPU_LIB=$L/libtamcmc_gpu_prev.so
done
