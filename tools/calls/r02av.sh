#!/bin/bash
set -x
mkdir -p gpurun_out
python bench.py --no-cpu > gpurun_out/r02av_bench_n1.json 2> gpurun_out/r02av_bench_n1.err
QC_PIPELINE_RUNNERS=1 python bench.py --no-cpu > gpurun_out/r02av_bench_n1_r1.json 2> gpurun_out/r02av_bench_n1_r1.err
echo done
