#!/bin/bash
# N-GPU check: the 2-rank gradient-equivalence test and a short bench line at N GPUs.  Usage: tools/gpu_ddp_check.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ddp_gpu.py tests/test_kernels_gpu.py -m gpu -q -k "two_rank or loss_finalize" --timeout 500 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2f_bench_base_${N}gpu.json 2> gpurun_out/err2.log || tail -8 gpurun_out/err2.log
cut -c1-330 gpurun_out/r2f_bench_base_${N}gpu.json
