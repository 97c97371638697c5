#!/bin/bash
# round 2, call t: 512-thread form at np = 1024, two runs per CTA (QC_GRID_VAR=2)
set -x
mkdir -p gpurun_out
timeout 300 python tools/variant_compare.py 1,2 ${1:-593} 16 1024 > gpurun_out/r02t_variants_1024.log 2>&1
echo done
