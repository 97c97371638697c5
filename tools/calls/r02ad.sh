#!/bin/bash
# round 2, call ad: np = 512 four-runs-per-CTA form with mbarrier hand-over
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py -x -q -m gpu -k "tensor_memory or ragged_krotov or packed_form or one_step_vs_oracle" > gpurun_out/r02ad_pytest_512.log 2>&1
echo "rc=$?" >> gpurun_out/r02ad_pytest_512.log
timeout 300 python tools/form_compare.py QC_GRID_TMP 512 0 > gpurun_out/r02ad_form512.log 2>&1
echo done
