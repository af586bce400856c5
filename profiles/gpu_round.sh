#!/bin/bash
# One GPU round (run under gpurun on ONE B200): parity tests, the bench line, the ncu launch list of the
# same bench command and one `--set full` capture of the fused kernel.  Outputs go to gpurun_out/<tag>_*;
# profiles/summarise_round.py turns them into the tracked summaries under profiles/.
#   usage: gpurun --timeout 1500 -- 'bash profiles/gpu_round.sh r1'
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
set -o pipefail
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err || { echo "bench failed"; tail -20 $OUT/${TAG}_bench.err; exit 1; }
cat $OUT/${TAG}_bench.json
python bench.py --impl reference --steps 10 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
cat $OUT/${TAG}_bench_reference.json
BCMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra"
$BCMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $BCMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:tamcmc_whittle_kernel -s 4 -c 2 -f -o $OUT/${TAG}_whittle $BCMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:tamcmc_expand_kernel -s 4 -c 2 -f -o $OUT/${TAG}_expand $BCMD > $OUT/${TAG}_ncu_expand.log 2>&1
echo "ncu expand rc=$?"
python profiles/bench_configs.py --configs c1,c4,c3,c5,env --steps 50 > $OUT/${TAG}_configs.jsonl 2> $OUT/${TAG}_configs.err
cat $OUT/${TAG}_configs.jsonl | cut -c1-200
# the red-giant set-up kernels (tamcmc_gpu_rgb_expand): launch list and one full capture of the two heavy ones, from the C4 config
RCMD="python profiles/bench_configs.py --configs c4 --steps 4"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tamcmc_rgb -c 40 --csv --log-file $OUT/${TAG}_rgb_launches.csv $RCMD > $OUT/${TAG}_ncu_rgb_launches.log 2>&1
echo "ncu rgb launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:tamcmc_rgb_pairs_kernel|tamcmc_rgb_search_kernel" -s 4 -c 2 -f -o $OUT/${TAG}_rgb $RCMD > $OUT/${TAG}_ncu_rgb.log 2>&1
echo "ncu rgb full rc=$?"
timeout 200 python profiles/l2_modes.py > $OUT/${TAG}_l2_modes.json 2> $OUT/${TAG}_l2_modes.err; cat $OUT/${TAG}_l2_modes.json
python profiles/far_accuracy.py > $OUT/${TAG}_far_accuracy.json 2> $OUT/${TAG}_far_accuracy.err; head -c 600 $OUT/${TAG}_far_accuracy.json
TAMCMC_GPU_LIB=$PWD/tamcmc-c_b200/libtamcmc_gpu_trace.so timeout 120 python profiles/trace_timeline.py > $OUT/${TAG}_trace.txt 2>&1; tail -14 $OUT/${TAG}_trace.txt | cut -c1-160
python __graft_entry__.py smoke 2>&1 | tail -2
ls -la $OUT | tail -14
