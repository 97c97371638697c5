#!/bin/bash
# round 2, call k: ncu full capture of the 512-thread form of kernel A
set -x
mkdir -p gpurun_out
python tools/variant_compare.py 2 296 4 > gpurun_out/r02k_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:grid8_sweep -s 2 -c 1 -o gpurun_out/r02k_kernelA_k8 -f python tools/variant_compare.py 2 296 4 > gpurun_out/r02k_ncu_stdout.log 2>&1
echo done
