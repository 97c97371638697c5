"""Time per step of small batches in the two forms of kernel A (one CTA per run / a cluster of two CTAs per run)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcontrol_b200 import control, engine, synthetic

ctx = engine.Context(0)
nt = 600
for np_, B in ((512, 1), (1024, 1), (2048, 1), (2048, 74), (2048, 148), (4096, 1), (4096, 74), (4096, 592)):
    pbs = synthetic.krotov_batch(B, np_, nt, seed=7)
    gb = control.GridBatch(pbs, ctx=ctx, store_traj=False)
    pl = gb.plan
    pl.set_E_base(gb._pad([gb._prescribed_field(p, control.FORWARD) for p in pbs]))
    out = {}
    for mode in ("0", "1"):
        if np_ == 4096 and mode == "0":
            continue
        os.environ["QC_GRID_CLUSTER"] = mode
        ts = []
        for rep in range(2):
            pl.snapshot_load(0)
            ctx.synchronize(); ctx.timer_start()
            pl.sweep(0, engine.FIELD_PRESCRIBED, 1, nt, True, -1, -1)
            ts.append(ctx.timer_stop())
        out[mode] = min(ts)
    print("np %5d  runs %4d : " % (np_, B) + "   ".join("%s %.1f us/step (%.3e pt*steps/s)" % (
        "cluster" if m == "1" else "one CTA", v * 1e3 / nt, B * np_ * nt / (v * 1e-3)) for m, v in out.items()), flush=True)
    gb.close()
