#!/bin/bash
# value / e2e of the full bench line against the number of batches in flight
for s in 4 5 6 8; do
  SPE_BENCH_SLOTS=$s python bench.py --steps 100 --warmup 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('slots', $s, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3), 'clock', d['clocks']['sm_mhz'], 'imgset', round(d['image_set']['images_per_s']))"
done
