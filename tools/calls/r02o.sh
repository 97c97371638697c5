#!/bin/bash
# round 2, call o: full GPU suite after the memory pool / arena / pipelined batch entry point
set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu -x > gpurun_out/r02o_pytest_full.log 2>&1
echo "rc=$?" >> gpurun_out/r02o_pytest_full.log
echo done
