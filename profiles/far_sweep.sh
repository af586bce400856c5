# A/B of the far-field folding on the bench workload (C2): look-ahead, ratio, number of terms
run() { tag=$1; shift; env "$@" python bench.py --no-cpu-baseline --steps 500 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$tag', round(d['value']), 'evals/s', round(d['ms_per_step']*1e3,1),'us/step kernel',round(r['kernel_ms']*1e3,1),'expand',round(r['expand_kernel_ms']*1e3,1),'frac',round(r['frac'],3),'e2e',round(d['e2e']['value']))"; }
run default X=1
run look2_end1 TAMCMC_GPU_LOOK=2 TAMCMC_GPU_LOOK_END=1
run ratio6 TAMCMC_GPU_FAR_RATIO=6
F20=$PWD/tamcmc-c_b200/libtamcmc_gpu_far20.so
if [ -f $F20 ]; then
run far20_ratio5 TAMCMC_GPU_LIB=$F20 TAMCMC_GPU_FAR_RATIO=5
run far20_ratio6 TAMCMC_GPU_LIB=$F20 TAMCMC_GPU_FAR_RATIO=6
fi
run nofar TAMCMC_GPU_FAR_RATIO=0
