#!/bin/bash
# round 2, call e: opt-in truncated series (test + per-step latency of the shipped configurations), kernel B after the retirement change
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edge.py -q -m gpu -k "truncated or ragged or unit_transform" > gpurun_out/r02e_pytest.log 2>&1
python tools/shipped_latency.py 2000 > gpurun_out/r02e_shipped_latency.log 2>&1
python tools/bench_nlevel.py --iters 12 --copies 1 20 > gpurun_out/r02e_bench_nlevel.log 2>&1
echo done
