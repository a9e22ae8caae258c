#!/bin/bash
# whole GPU suite, then the quick bench with and without the decoder layer-0 fold
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for f in 1 0 1 0; do
  SPE_DEC0_FOLD=$f python bench.py --quick --steps 60 --warmup 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dec0 fold', $f, 'ms', round(d['ms_per_step'],3), 'solved', d['poses_solved_per_batch'], d['gpu_launches_by_family'])"
done
