#!/bin/bash
# accuracy + speed of the decoder / kv 3xTF32 switches
for cfg in "1 1" "1 0" "0 0"; do
  set -- $cfg
  echo "== SPE_DEC_X3=$1 SPE_KV_X3=$2"
  SPE_DEC_X3=$1 SPE_KV_X3=$2 python -m pytest tests/test_gpu_bench_configs.py -m gpu -q -s -k "random_init or whole_chain" 2>&1 | grep -E "random-init|chain B|chain poses|passed|failed"
  SPE_DEC_X3=$1 SPE_KV_X3=$2 python bench.py --quick --steps 40 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('quick', round(d['value']), round(d['ms_per_step'],3))"
done
