#!/bin/bash
# round 2, call c: full GPU suite, the default bench line with the new blocks, the reference arm, launch list + DRAM bytes
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02c_pytest_full.log 2>&1
python bench.py > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02c_bench_ref.json 2> gpurun_out/r02c_bench_ref.err
python bench.py --no-cpu --no-extra --steps 1 --warmup 3 > gpurun_out/r02c_bench_plain.json 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02c_launches.csv python bench.py --no-cpu --no-extra --steps 1 --warmup 3 > gpurun_out/r02c_ncu_stdout.log 2>&1
echo done
