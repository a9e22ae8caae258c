#!/bin/bash
# parity at the benchmarked configs, then the quick bench a few times
python -m pytest tests/test_gpu_bench_configs.py -m gpu -q -s 2>&1 | grep -E "random-init|chain B|B=256|passed|failed|^E "
for i in 1 2 3; do
  python bench.py --quick --steps 60 --warmup 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('quick ms', round(d['ms_per_step'],3), 'solved', d['poses_solved_per_batch'])"
done
