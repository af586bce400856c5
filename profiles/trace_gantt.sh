# usage (on the GPU box): bash profiles/trace_gantt.sh [stagger_ns]
#   build first: make -C tamcmc-c_b200/csrc EXTRA="-DTAMCMC_TRACE -DTAMCMC_TRACE_GANTT" OUT=../libtamcmc_gpu_gantt.so B=build_gantt
TAMCMC_GPU_STAGGER_NS=${1:-0} TAMCMC_GPU_LIB=$PWD/tamcmc-c_b200/libtamcmc_gpu_gantt.so timeout 120 python profiles/trace_gantt.py
