#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ddp_gpu.py -m gpu -q --timeout 500 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2e_bench_base_2gpu.json 2> gpurun_out/err2.log || tail -8 gpurun_out/err2.log
cut -c1-330 gpurun_out/r2e_bench_base_2gpu.json
