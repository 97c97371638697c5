#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/e2e_api_probe.py 1184 1024 > gpurun_out/r02am_e2e_api.log 2>&1
timeout 600 python -m pytest tests/test_bench_contract.py -q -m gpu > gpurun_out/r02am_pytest.log 2>&1
echo done
