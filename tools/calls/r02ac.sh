#!/bin/bash
# round 2, call ac: cluster form with the late arrive -- latency mode table, size sweep, cluster tests
set -x
mkdir -p gpurun_out
timeout 300 python tools/latency_mode.py > gpurun_out/r02ac_latency_mode.log 2>&1
timeout 300 python tools/size_sweep.py > gpurun_out/r02ac_size_sweep.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu -k "4096 or cluster or latency or edge" > gpurun_out/r02ac_pytest.log 2>&1
echo done
