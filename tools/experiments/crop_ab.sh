#!/bin/bash
# staged vs gather crop kernel (CUDA events), then ncu --set full of the staged one
python tests/probes/ncu_probe_r02.py 6 2>&1 | grep crop
SPE_CROP_GATHER=1 python tests/probes/ncu_probe_r02.py 6 2>&1 | grep crop
ncu --set full --clock-control none --import-source on -k regex:crop_resize -c 3 -o gpurun_out/r02d_crop python tests/probes/ncu_probe_r02.py 3 > gpurun_out/r02d_crop_ncu.log 2>&1
tail -2 gpurun_out/r02d_crop_ncu.log
