#!/bin/bash
python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "second_operand" 2>&1 | tail -5
python -m pytest tests/test_gpu_crop.py -m gpu -q -x 2>&1 | tail -2
SPE_FUSE_DOWN=1 python -m pytest tests/test_gpu_bench_configs.py -m gpu -q -s 2>&1 | grep -E "random-init|chain B|B=256|passed|failed|^E "
for f in 1 0 1 0; do
  SPE_FUSE_DOWN=$f python bench.py --quick --steps 60 --warmup 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fuse_down', $f, 'ms', round(d['ms_per_step'],3), 'solved', d['poses_solved_per_batch'])"
done
python tests/probes/ncu_probe_r02.py 6 2>&1 | grep crop | tail -3
