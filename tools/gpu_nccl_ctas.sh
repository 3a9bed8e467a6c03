#!/bin/bash
# N-GPU A/B of NCCL's CTA budget: its all-reduce kernels take SMs from the persistent GEMMs of the backward
N=${1:-2}
port=29520
for c in default 8 2; do
  port=$((port + 1))
  if [ "$c" != default ]; then export NCCL_MAX_CTAS=$c; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 30 --warmup 5 --quick > gpurun_out/nccl_$c.json 2> gpurun_out/nccl_$c.err || tail -5 gpurun_out/nccl_$c.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/nccl_$c.json').read().strip().splitlines()[-1]); print('NCCL_MAX_CTAS=$c', round(d['ms_step'],3), round(d['images_per_s']))"
done
