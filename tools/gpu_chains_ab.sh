#!/bin/bash
# round 2e: parity of the row-chain engine + A/B of chains x PDL on the three bench workloads
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2e_pytest_gpu.txt 2>&1
tail -15 gpurun_out/r2e_pytest_gpu.txt
: > gpurun_out/r2e_ab.jsonl
for cfg in "1 1" "2 1" "2 0" "1 0"; do
  set -- $cfg
  CSMAE_CHAINS=$1 CSMAE_PDL=$2 timeout 300 python bench.py --quick --steps 30 --warmup 5 >> gpurun_out/r2e_ab.jsonl 2>gpurun_out/err.log || tail -5 gpurun_out/err.log
done
for c in 1 2; do
  CSMAE_CHAINS=$c timeout 300 python bench.py --quick --arch large --steps 20 --warmup 5 >> gpurun_out/r2e_ab.jsonl 2>gpurun_out/err.log || tail -5 gpurun_out/err.log
  CSMAE_CHAINS=$c timeout 300 python bench.py --quick --arch large --batch 16 --input-size 448 --steps 15 --warmup 5 >> gpurun_out/r2e_ab.jsonl 2>gpurun_out/err.log || tail -5 gpurun_out/err.log
done
cat gpurun_out/r2e_ab.jsonl
