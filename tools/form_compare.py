"""Kernel A, default form against an alternative selected by an environment switch:

    python tools/form_compare.py SWITCH [sizes] [batches]      e.g.  QC_GRID_PACK 512,256 0,7,1

SWITCH is QC_GRID_PACK or QC_GRID_CLUSTER; batch 0 = a batch that fills the GPU.  Prescribed-field sweep of 16
steps per measurement; prints the throughput of either form and the largest L2 distance between their results."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcontrol_b200 import control, engine, synthetic

switch = sys.argv[1] if len(sys.argv) > 1 else "QC_GRID_PACK"
sizes = [int(s) for s in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2048, 1024, 512, 256]
batches = [int(s) for s in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
ctx = engine.Context(0)
peak = ctx.fp64_peak_tflops()
nt = 16


def compare(np_, B):
    pbs = synthetic.krotov_batch(B, np_, nt, seed=3)
    gb = control.GridBatch(pbs, ctx=ctx, store_traj=False)
    pl = gb.plan
    pl.set_E_base(gb._pad([gb._prescribed_field(p, control.FORWARD) for p in pbs]))
    out = {}
    for form in ("0", "1", "0", "1"):
        os.environ[switch] = form
        ts = []
        for rep in range(3):
            pl.snapshot_load(0)
            ctx.synchronize(); ctx.timer_start()
            pl.sweep(0, engine.FIELD_PRESCRIBED, 1, nt, True, -1, -1)
            ts.append(ctx.timer_stop())
        ms = min(ts[1:])
        out[form] = pl.get_state()
        flops = ((64 - 1) * 2 * np_ * (10 * int(np.log2(np_)) + 18) + 18 * 2 * np_) * B * nt
        print(json.dumps({"np": np_, "runs": B, switch: form, "ms": ms,
                          "grid_pt_timesteps_per_s": B * np_ * nt / (ms * 1e-3),
                          "frac_of_measured_fp64_peak": flops / (ms * 1e-3) / 1e12 / peak}), flush=True)
    d = out["0"] - out["1"]
    print(json.dumps({"np": np_, "runs": B, "max_l2_difference_between_forms_after_%d_steps" % nt:
                      float(np.sqrt(np.max(np.sum(np.abs(d) ** 2, axis=(1, 2, 3)) * pbs[0].dx)))}), flush=True)
    gb.close()


for np_ in sizes:
    for B in batches:
        compare(np_, B or 148 * (28 if np_ >= 1024 else 56))
