#!/bin/bash
# round 2, call b: new GPU tests (stop rule, >2 GiB pitch, memory chunks, thread pool), ncu full capture of kernel A with the TMEM hand-over
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edge.py tests/test_gpu_dropin.py -q -m gpu -k "stopping_rule or 2gib or memory_sized or thread_pool or fitting_solver_drop_in or batch_entry or newcheb_command" > gpurun_out/r02b_pytest_new.log 2>&1
python tools/variant_compare.py 1 296 4 > gpurun_out/r02b_small_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:grid_sweep -s 2 -c 2 -o gpurun_out/r02b_kernelA_tmh -f python tools/variant_compare.py 1 296 4 > gpurun_out/r02b_ncu_stdout.log 2>&1
echo done
