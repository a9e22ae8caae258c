#!/bin/bash
python -m pytest tests/test_gpu_bench_configs.py tests/test_gpu_model.py -m gpu -q -s 2>&1 | grep -E "random-init|chain B|chain poses|B=256|passed|failed|^E "
for f in 0 1 0 1; do
  SPE_FOLD_NECK=$f python bench.py --quick --steps 60 --warmup 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fold', $f, 'ms', round(d['ms_per_step'],3), 'solved', d['poses_solved_per_batch'])"
done
