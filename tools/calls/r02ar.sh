#!/bin/bash
# round 2, call ar: the reference's shipped configurations at FULL length with the final build (cluster form with the
# late arrive and split pushes is what a lone np = 2048 run uses)
set -x
mkdir -p gpurun_out
QC_SHIPPED_FULL=1 timeout 1500 python -m pytest tests/test_gpu_shipped.py -q -m gpu -s > gpurun_out/r02ar_shipped_full.log 2>&1
echo "rc=$?" >> gpurun_out/r02ar_shipped_full.log
echo done
