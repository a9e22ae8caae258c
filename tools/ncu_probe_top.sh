#!/bin/bash
# ncu --set full capture of the fused feed-forward kernel (run the probe plainly first)
set -e
python tools/ffn_probe.py 3 > gpurun_out/ffn_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ffn_tc --launch-skip 2 -c 1 -o gpurun_out/r01i_ffn -f python tools/ffn_probe.py 3 > gpurun_out/ncu_ffn.log 2>&1
