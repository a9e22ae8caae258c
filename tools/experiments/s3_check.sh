#!/bin/bash
# whole GPU suite, SA forward timing, bf16 batch-256 side config with and without the fused feed-forward
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/sa_probe.py 64
for v in 1 0; do
  SPE_MIXED_FFN=$v SPE_BENCH_BATCH=256 SPE_BENCH_PRECISION=bf16 SPE_BENCH_SLOTS=3 python bench.py --quick --steps 12 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bf16 b256 mixed_ffn=$v', round(d['value']), round(d['ms_per_step'],3), d['poses_solved_per_batch'])"
done
