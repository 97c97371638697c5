#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/variant_compare.py 1 148 4 4096 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:grid_sweep -s 2 -c 1 -o gpurun_out/r02af_kernelA_4096 -f python tools/variant_compare.py 1 148 4 4096 > gpurun_out/r02af_ncu_stdout.log 2>&1
echo done
