#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "512_thread_form or bench_workload" > gpurun_out/r02w_pytest.log 2>&1
echo done
