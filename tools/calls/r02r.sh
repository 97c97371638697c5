#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edge.py -q -m gpu -k "pool_and_arena or memory_sized" > gpurun_out/r02r_pytest_pool.log 2>&1
timeout 600 python tools/e2e_api_probe.py 1024 1024 > gpurun_out/r02r_e2e_api.log 2>&1
echo done
