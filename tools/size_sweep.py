"""Throughput of kernel A at every grid size with a batch that fills the GPU (prescribed-field sweep, 16 steps)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcontrol_b200 import control, engine, synthetic

ctx = engine.Context(0)
peak = ctx.fp64_peak_tflops()
nt = 16
for np_, B in ((256, 148 * 64), (512, 148 * 56), (1024, 148 * 56), (2048, 148 * 28), (4096, 74 * 16)):
    pbs = synthetic.krotov_batch(B, np_, nt, seed=3)
    gb = control.GridBatch(pbs, ctx=ctx, store_traj=False)
    pl = gb.plan
    pl.set_E_base(gb._pad([gb._prescribed_field(p, control.FORWARD) for p in pbs]))
    ts = []
    for rep in range(3):
        pl.snapshot_load(0)
        ctx.synchronize(); ctx.timer_start()
        pl.sweep(0, engine.FIELD_PRESCRIBED, 1, nt, True, -1, -1)
        ts.append(ctx.timer_stop())
    ms = min(ts[1:])
    log2 = int(np.log2(np_))
    flops = ((64 - 1) * 2 * np_ * (10 * log2 + 18) + 18 * 2 * np_) * B * nt
    print(json.dumps({"np": np_, "runs": B, "ms": ms, "grid_pt_timesteps_per_s": B * np_ * nt / (ms * 1e-3),
                      "tflops_algorithmic": flops / (ms * 1e-3) / 1e12, "frac_of_measured_fp64_peak": flops / (ms * 1e-3) / 1e12 / peak}), flush=True)
    gb.close()
