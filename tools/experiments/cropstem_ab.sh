#!/bin/bash
# pipeline tests, then the quick bench with the crop writing the stem input directly (default) and without
python -m pytest tests/test_gpu_model.py tests/test_gpu_bench_configs.py tests/test_gpu_crop.py -m gpu -x -q 2>&1 | tail -4
for v in 1 0 1 0; do
  SPE_CROP_STEM=$v python bench.py --quick --steps 80 --warmup 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('crop->stem', $v, 'ms', round(d['ms_per_step'],3), 'solved', d['poses_solved_per_batch'], d['gpu_launches_by_family'])"
done
