#!/bin/bash
set -x
mkdir -p gpurun_out
for rep in 1 2; do
QCONTROL_B200_LIB=$PWD/qcontrol_b200/lib/alt_cl2.so timeout 300 python tools/variant_compare.py 1 1184 16 4096 >> gpurun_out/r02ab_cl2.log 2>&1
timeout 300 python tools/variant_compare.py 1 1184 16 4096 >> gpurun_out/r02ab_cl1.log 2>&1
done
python tools/variant_compare.py 1 148 4 4096 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:grid_sweep -s 2 -c 1 -o gpurun_out/r02ab_kernelA_4096 -f python tools/variant_compare.py 1 148 4 4096 > gpurun_out/r02ab_ncu_stdout.log 2>&1
echo done
