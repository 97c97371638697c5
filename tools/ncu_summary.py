"""Summarise one kernel of an .ncu-rep (ncu --set full) as `metric,unit,value` rows for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "# comment line" ... > profiles/rNN_summary.csv
"""
import csv, io, subprocess, sys

METRICS = """gpu__time_duration.sum launch__registers_per_thread launch__block_size launch__grid_size
launch__shared_mem_per_block_dynamic sm__warps_active.avg.pct_of_peak_sustained_active
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active l1tex__throughput.avg.pct_of_peak_sustained_active
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active sm__cycles_elapsed.max
dram__bytes_read.sum dram__bytes_write.sum smsp__sass_inst_executed_op_local_ld.sum
smsp__sass_inst_executed_op_local_st.sum smsp__sass_inst_executed_op_tmem_ldt.sum
smsp__sass_inst_executed_op_tmem_stt.sum smsp__sass_thread_inst_executed_op_dfma_pred_on.sum
smsp__sass_thread_inst_executed_op_dadd_pred_on.sum smsp__sass_thread_inst_executed_op_dmul_pred_on.sum
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio""".split()

rep = sys.argv[1]
for c in sys.argv[2:]:
    print(c)
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
print("# kernel: %s" % vals[hdr.index("Kernel Name")])
print("metric,unit,value")
for m in METRICS:
    if m in hdr:
        i = hdr.index(m)
        print("%s,%s,%s" % (m, units[i], vals[i].replace(",", "")))
