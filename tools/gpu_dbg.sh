#!/bin/bash
for rep in 1 2; do
for cfg in "CSMAE_PDL=1" "CSMAE_PDL=0" "CSMAE_SIDE_STREAM=0"; do
  env $cfg timeout 300 python bench.py --quick --steps 40 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', round(d['ms_step'],3), d['clocks']['sm_mhz'])"
done; done
