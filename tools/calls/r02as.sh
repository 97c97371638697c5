#!/bin/bash
# round 2, call as (2 GPUs): FINAL commit -- the whole GPU suite including the 2-rank NCCL test, smoke()
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02as_smoke.log 2>&1
timeout 1700 python -m pytest tests -q -m gpu > gpurun_out/r02as_pytest_full.log 2>&1
echo "rc=$?" >> gpurun_out/r02as_pytest_full.log
echo done
