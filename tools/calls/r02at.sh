#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/e2e_api_probe.py 1184 1024 > gpurun_out/r02at_e2e_api.log 2>&1
QC_PIPELINE_RUNNERS=1 timeout 600 python tools/e2e_api_probe.py 1184 1024 > gpurun_out/r02at_e2e_api_r1.log 2>&1
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_edge.py -x -q -m gpu > gpurun_out/r02at_pytest.log 2>&1
echo done
