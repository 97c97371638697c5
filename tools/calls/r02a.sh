#!/bin/bash
# round 2, call a: TMEM hand-over variant of kernel A vs the round-1 kernel, parity under the variant, FP64 op counts, full GPU suite
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt
python tools/variant_compare.py 0,1,0,1 4144 16 > gpurun_out/r02a_variants.log 2> gpurun_out/r02a_variants.err
QC_GRID_VAR=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py -x -q -m gpu -k "one_step or full_size or propagate_host or krotov or bench_workload" > gpurun_out/r02a_pytest_var1.log 2>&1
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02a_pytest_full.log 2>&1
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum,sm__cycles_elapsed.max,sm__cycles_active.avg,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_fp64_pred_on.sum
python tools/variant_compare.py 0,1 296 4 > gpurun_out/r02a_small_plain.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:grid_sweep --csv --log-file gpurun_out/r02a_opcounts.csv python tools/variant_compare.py 0,1 296 4 > gpurun_out/r02a_ncu_stdout.log 2>&1
echo done
