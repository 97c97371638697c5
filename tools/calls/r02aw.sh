#!/bin/bash
# round 2, call aw: whole GPU suite and smoke() on the final commit (one GPU)
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02aw_smoke.log 2>&1
timeout 1700 python -m pytest tests -q -m gpu > gpurun_out/r02aw_pytest_full.log 2>&1
echo "rc=$?" >> gpurun_out/r02aw_pytest_full.log
echo done
