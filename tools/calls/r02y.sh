#!/bin/bash
# round 2, call y: final build -- full GPU suite, default bench line + reference arm, launch list with DRAM bytes, ncu full capture of the Krotov forward sweep
set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu > gpurun_out/r02y_pytest_full.log 2>&1
echo "rc=$?" >> gpurun_out/r02y_pytest_full.log
python bench.py > gpurun_out/r02y_bench_n1.json 2> gpurun_out/r02y_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02y_bench_ref.json 2> gpurun_out/r02y_bench_ref.err
python bench.py --no-cpu --no-extra --steps 1 --warmup 3 > gpurun_out/r02y_bench_plain.json 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02y_launches.csv python bench.py --no-cpu --no-extra --steps 1 --warmup 3 > gpurun_out/r02y_ncu_stdout.log 2>&1
python tools/variant_compare.py 1 296 4 > gpurun_out/r02y_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:grid_sweep -s 8 -c 1 -o gpurun_out/r02y_kernelA_final -f python tools/variant_compare.py 1 296 4 > gpurun_out/r02y_ncu2_stdout.log 2>&1
echo done
