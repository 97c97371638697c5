#!/bin/bash
# round 2, call f: TMA-staged trajectory slice (Krotov sweeps at np = 2048), truncated series, shipped-configuration latency
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bench_workload or tensor_memory or full_size" > gpurun_out/r02f_pytest_tma.log 2>&1
echo "rc=$?" >> gpurun_out/r02f_pytest_tma.log
timeout 600 python -m pytest tests/test_gpu_edge.py -q -m gpu -k "truncated or stopping" > gpurun_out/r02f_pytest_trunc.log 2>&1
timeout 300 python tools/shipped_latency.py 2000 > gpurun_out/r02f_shipped_latency.log 2>&1
timeout 300 python tools/variant_compare.py 1,0 4144 16 > gpurun_out/r02f_variants.log 2>&1
timeout 300 python tools/bench_nlevel.py --iters 12 --copies 1 20 > gpurun_out/r02f_bench_nlevel.log 2>&1
echo done
