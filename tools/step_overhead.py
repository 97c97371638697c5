"""Kernel A: time of a step as a function of the polynomial order, T(nch) = a + (nch - 1) b, to separate the
per-step work outside the order loop (rotating frame, field, norms, renormalisation, trajectory I/O) from the loop."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcontrol_b200 import control, engine, synthetic

ctx = engine.Context(0)
nt = 16
for np_, B, waves in ((2048, 148 * 28, 28), (1024, 148 * 56, 28), (256, 148 * 64, 8)):
    res = {}
    for nch in (16, 32, 64, 128):
        pbs = synthetic.krotov_batch(B, np_, nt, nch, seed=3)
        gb = control.GridBatch(pbs, ctx=ctx, store_traj=False)
        pl = gb.plan
        pl.set_E_base(gb._pad([gb._prescribed_field(p, control.FORWARD) for p in pbs]))
        ts = []
        for rep in range(3):
            pl.snapshot_load(0)
            ctx.synchronize(); ctx.timer_start()
            pl.sweep(0, engine.FIELD_PRESCRIBED, 1, nt, True, -1, -1)
            ts.append(ctx.timer_stop())
        res[nch] = min(ts[1:]) * 1e3 / (waves * nt)          # us per step of one resident wave
        gb.close()
    b = (res[128] - res[32]) / 96.0
    a = res[64] - 63 * b
    print(json.dumps({"np": np_, "us_per_step": res, "us_per_order": b, "us_outside_the_order_loop": a,
                      "share_at_nch_64": a / res[64]}), flush=True)
