#!/bin/bash
# launch list of the SA predictor's forward (no graph => launch order = schedule order); only after the plain run exited 0
TAG=${1:-r02f_sa}
export SPE_NO_GRAPH=1
python tools/sa_probe.py 64 > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; cat gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -c 1700 --csv --log-file gpurun_out/launches_$TAG.csv python tools/sa_probe.py 64 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log; cat gpurun_out/plain_$TAG.log
