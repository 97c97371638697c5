#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bench_workload" > gpurun_out/r02l_pytest_k8.log 2>&1
timeout 300 python tools/variant_compare.py 1,2 4144 16 > gpurun_out/r02l_variants.log 2>&1
echo done
