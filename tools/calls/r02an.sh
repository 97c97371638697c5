#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_edge.py -x -q -m gpu > gpurun_out/r02an_pytest.log 2>&1
echo done
