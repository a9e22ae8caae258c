#!/bin/bash
python -m pytest tests/test_gpu_model.py tests/test_gpu_bench_configs.py -m gpu -x -q 2>&1 | tail -3
for v in 1 0 1 0; do
  SPE_LN_ROWS4=$v python bench.py --quick --steps 80 --warmup 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ln rows4', $v, 'ms', round(d['ms_per_step'],3))"
done
