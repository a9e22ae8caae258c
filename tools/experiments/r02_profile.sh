#!/bin/bash
# launch list of one step (one batch in flight, no graph => launch order = schedule order) with per-launch duration,
# DRAM bytes and tensor-pipe activity; only after the same command has exited 0 without ncu
TAG=${1:-r02b}
export SPE_BENCH_SLOTS=1 SPE_NO_GRAPH=1
python bench.py --quick --steps 2 --warmup 3 > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -c 1400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
