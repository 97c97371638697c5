"""CUDA-event time of each sweep kind on the bench workload (debugging aid for bench.py's roofline leg)."""
import os, sys, cmath
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcontrol_b200 import control, engine, synthetic

B, nt = int(os.environ.get("RUNS", 4144)), 16
pbs = synthetic.krotov_batch(B, 2048, nt, seed=12345)
ctx = engine.Context(0)
gb = control.GridBatch(pbs, ctx=ctx)
pl = gb.plan
gb.time_sweeps = False
hist = gb.run_krotov(iter_max=1)           # leaves both trajectory buffers filled
sq = np.array([np.complex128(cmath.sqrt(p.laser_field_hf(np.float64(1.0), p.T, p.pcos, p.w_list))) for p in pbs])

def timed(label, fn):
    ctx.synchronize(); ctx.timer_start(); fn(); ms = ctx.timer_stop()
    print("%-28s %8.2f ms   %.3e grid-pt*steps/s" % (label, ms, B * 2048 * nt / (ms * 1e-3)))

for rep in range(2):
    pl.snapshot_load(0)
    timed("prescribed fwd (no traj)", lambda: pl.sweep(0, engine.FIELD_PRESCRIBED, 1, nt, True, -1, -1))
    pl.snapshot_load(0)
    timed("prescribed fwd (write traj)", lambda: pl.sweep(0, engine.FIELD_PRESCRIBED, 1, nt, True, -1, 0))
    pl.snapshot_load(0)
    timed("krotov fwd", lambda: pl.sweep(0, engine.FIELD_KROTOV_FWD, 1, nt, True, 1, 0))
    pl.krotov_terminal(1, sq, 1)
    timed("krotov bwd", lambda: pl.sweep(1, engine.FIELD_KROTOV_BWD, 1, nt, True, 0, 1))
    pl.snapshot_load(0)
    timed("prescribed bwd-slot (no traj)", lambda: pl.sweep(1, engine.FIELD_PRESCRIBED, 1, nt, True, -1, -1))

print("--- the bench.py iteration, piece by piece, then as a whole")
import torch
def iteration(timed_pieces):
    t = timed if timed_pieces else (lambda label, fn: fn())
    pl.snapshot_load(0)
    pl.record_state(0, 0, 0)
    t("  fwd (l=0..nt)", lambda: pl.sweep(0, engine.FIELD_KROTOV_FWD, 0, nt + 1, True, 1, 0))
    t("  goal overlap", lambda: pl.goal_overlap(1))
    t("  terminal", lambda: pl.krotov_terminal(1, sq, 1))
    t("  bwd", lambda: pl.sweep(1, engine.FIELD_KROTOV_BWD, 1, nt, True, 0, 1))
for rep in range(2):
    iteration(True)
for rep in range(2):
    ctx.synchronize(); ctx.timer_start()
    for k in range(3):
        iteration(False)
    ms = ctx.timer_stop()
    print("3 whole iterations: %.2f ms each" % (ms / 3))
