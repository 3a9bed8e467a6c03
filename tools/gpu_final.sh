#!/bin/bash
# last pass of a round on the final code: GPU tests, smoke, the bench line and the reference arm's line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2g_pytest_gpu.txt; cat gpurun_out/r2g_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2g_bench_base.json 2> gpurun_out/r2g_bench_base.err; cut -c1-200 gpurun_out/r2g_bench_base.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2g_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/r2g_bench_reference.json
