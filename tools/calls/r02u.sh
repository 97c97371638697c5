#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/variant_compare.py 2 592 4 1024 > gpurun_out/r02u_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:grid8x2 -s 2 -c 1 -o gpurun_out/r02u_kernelA_k8x2 -f python tools/variant_compare.py 2 592 4 1024 > gpurun_out/r02u_ncu_stdout.log 2>&1
echo done
