TAMCMC_GPU_LIB=$PWD/tamcmc-c_b200/libtamcmc_gpu_trace.so timeout 120 python profiles/trace_timeline.py "$@" 2>&1 | grep "^cta" | sed 's/np.int64(\([-0-9]*\))/\1/g' | cut -c1-420
