#!/bin/bash
# effective cost of each stage with 4 batches in flight: step time with the stage left out
for m in 0 1 2 4 8 16 32 64 128 256 0; do
  SPE_DBG_SKIP=$m python bench.py --quick --steps 60 --warmup 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('skip', $m, 'ms', round(d['ms_per_step'],3))"
done
