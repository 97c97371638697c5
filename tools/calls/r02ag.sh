#!/bin/bash
# round 2, call ag: final build -- smoke(), full GPU suite, default bench line + reference arm
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ag_smoke.log 2>&1
timeout 1700 python -m pytest tests -q -m gpu > gpurun_out/r02ag_pytest_full.log 2>&1
echo "rc=$?" >> gpurun_out/r02ag_pytest_full.log
python bench.py > gpurun_out/r02ag_bench_n1.json 2> gpurun_out/r02ag_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02ag_bench_ref.json 2> gpurun_out/r02ag_bench_ref.err
echo done
