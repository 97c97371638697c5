#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/size_sweep.py > gpurun_out/r02aa_size_sweep.log 2>&1
timeout 600 python -m pytest tests -x -q -m gpu -k "4096 or cluster" > gpurun_out/r02aa_pytest.log 2>&1
echo done
