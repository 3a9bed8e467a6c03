#!/bin/bash
# One GPU-box pass that produces the per-round evidence under gpurun_out/ (copied to profiles/ afterwards):
#   tests, DRAM traffic of every GEMM launch of one step (written where bench.py looks for it), the bench line, the ncu
#   launch list of the bench command and of one graphed step, and `--set full` summaries of the top kernels.
#   Usage: tools/gpu_profile_round.sh <tag>
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q 2>&1 | tail -5 > $OUT/${TAG}_pytest_gpu.txt
tail -2 $OUT/${TAG}_pytest_gpu.txt
# DRAM traffic of every GEMM launch of ONE un-graphed step (one pass per kernel); bench.py checks the build identity
timeout 600 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --clock-control none -k regex:gemm_kernel --csv --log-file $OUT/${TAG}_gemm_traffic.csv python tools/step_for_ncu.py > /dev/null 2>&1
python tools/ncu_traffic.py $OUT/${TAG}_gemm_traffic.csv profiles/r2_gemm_dram_traffic.json > /dev/null 2>&1
cp profiles/r2_gemm_dram_traffic.json $OUT/r2_gemm_dram_traffic.json
python bench.py > $OUT/${TAG}_bench_base.json 2> $OUT/${TAG}_bench_base.err
cut -c1-400 $OUT/${TAG}_bench_base.json
# launch list of the bench command (graph kernel nodes are profiled one by one) and of one graphed step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_launches_bench_base.csv \
    python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
python tools/ncu_launch_summary.py $OUT/${TAG}_launches_bench_base.csv > $OUT/${TAG}_launches_summary.txt 2>&1
CSM_NCU_GRAPHS=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_launches_graph_step.csv python tools/step_for_ncu.py > /dev/null 2>&1
python tools/ncu_launch_summary.py $OUT/${TAG}_launches_graph_step.csv > $OUT/${TAG}_launches_graph_step_summary.txt 2>&1
head -14 $OUT/${TAG}_launches_graph_step_summary.txt
# full-set captures of a few launches of the top kernel families, summarised here (the reports stay on the box)
cap() {  # name, kernel regex, skip, count
  timeout 600 ncu --profile-from-start off --set full --clock-control none -k "regex:$2" --launch-skip $3 -c $4 \
      -o $OUT/${TAG}_full_$1 python tools/step_for_ncu.py > $OUT/${TAG}_full_$1.log 2>&1
  python tools/ncu_summary.py $OUT/${TAG}_full_$1.ncu-rep $OUT/${TAG}_ncu_full_$1_summary.csv > /dev/null 2>&1
  rm -f $OUT/${TAG}_full_$1.ncu-rep
}
cap gemm gemm_kernel 1 12
cap layernorm layernorm 38 8
cap bn_patch bn_patch 0 2
cap attention attn_ 0 40
ls -la $OUT | tail -24
