#!/bin/bash
python -m pytest tests/test_gpu_kernels.py -m gpu -q -k attention 2>&1 | tail -3
for safe in 1 0 1 0; do
  SPE_ATTN_SAFE=$safe python bench.py --quick --steps 60 --warmup 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('attn safe', $safe, 'ms', round(d['ms_per_step'],3), 'solved', d['poses_solved_per_batch'])"
done
python -m pytest tests/test_gpu_bench_configs.py tests/test_gpu_model.py -m gpu -q 2>&1 | tail -3
