#!/bin/bash
# accuracy + speed of the plain-TF32 decoder feed-forward (SPE_DEC_FFN_X3=0) against the 3xTF32 default
for v in 1 0 1 0; do
  echo "== SPE_DEC_FFN_X3=$v"
  SPE_DEC_FFN_X3=$v python -m pytest tests/test_gpu_bench_configs.py -m gpu -q -s -k "random_init or whole_chain or batch256" 2>&1 | grep -E "random-init|chain B|chain poses|B=256|passed|failed|^E "
  SPE_DEC_FFN_X3=$v python bench.py --quick --steps 60 --warmup 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('quick', round(d['value']), round(d['ms_per_step'],3))"
done
