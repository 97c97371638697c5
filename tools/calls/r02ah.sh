#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dropin.py -x -q -m gpu -k "pipelined_pieces or batch_entry" > gpurun_out/r02ah_pytest.log 2>&1
echo done
