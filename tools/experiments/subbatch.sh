#!/bin/bash
for cfg in "0 4" "32 4" "32 2" "16 4" "16 2" "32 1" "0 2" "0 6"; do
  set -- $cfg
  SPE_SUBBATCH=$1 SPE_BENCH_SLOTS=$2 python bench.py --quick --steps 60 --warmup 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('subbatch', $1, 'slots', $2, 'ms', round(d['ms_per_step'],3))"
done
