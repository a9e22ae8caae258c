#!/bin/bash
# stem + max-pool + layer1 in L2-resident image chunks (SPE_HEAD_CHUNK): parity at the benchmarked configs, then speed
SPE_HEAD_CHUNK=16 python -m pytest tests/test_gpu_bench_configs.py -m gpu -q -s 2>&1 | grep -E "random-init|chain B|B=256|passed|failed|^E "
for hc in 0 16 32 8 0 16; do
  SPE_HEAD_CHUNK=$hc python bench.py --quick --steps 60 --warmup 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('head_chunk', $hc, 'ms', round(d['ms_per_step'],3), 'solved', d['poses_solved_per_batch'])"
done
