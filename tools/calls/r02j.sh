#!/bin/bash
# round 2, call j: first run of the 512-thread form of kernel A (QC_GRID_VAR=2)
set -x
mkdir -p gpurun_out
timeout 300 python tools/variant_compare.py 1,2 ${1:-4144} 16 > gpurun_out/r02j_variants.log 2>&1
echo done
