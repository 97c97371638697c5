#!/bin/bash
# round 2, call m: pipelined batch entry point (set-up of piece k + 1 beside the control loops of piece k)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_edge.py -x -q -m gpu -k "batch_entry or newcheb_command or memory_sized or thread_pool" > gpurun_out/r02m_pytest.log 2>&1
timeout 600 python tools/e2e_api_probe.py 1024 1024 > gpurun_out/r02m_e2e_api.log 2>&1
timeout 300 python tools/e2e_api_probe.py 1024 256 >> gpurun_out/r02m_e2e_api.log 2>&1
echo done
