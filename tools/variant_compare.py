"""Kernel A at np = 2048 (plain one-run-per-CTA form): the kernel variants selected by QC_GRID_VAR, side by side.

    python tools/variant_compare.py [variants] [runs] [nt]        e.g.  0,1 4144 16

For every variant: (a) a prescribed-field sweep and (b) a Krotov iteration (forward sweep with the in-kernel field
update from the stored backward trajectory + backward sweep) of a batch that fills the GPU, each timed with the
library's CUDA events on the launching stream (best of 3 after a warm-up), the fraction of the FP64 peak measured in
the same process, and the L2 distance of the final states and of the optimised fields from variant 0."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcontrol_b200 import control, engine, synthetic

variants = [int(s) for s in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 28
nt = int(sys.argv[3]) if len(sys.argv) > 3 else 16
np_ = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
ctx = engine.Context(0)
peak = ctx.fp64_peak_tflops()
log2 = int(np.log2(np_))
F_presc = ((64 - 1) * 2 * np_ * (10 * log2 + 18) + 18 * 2 * np_) * B * nt
F_krot = F_presc + 8 * np_ * B * nt

pbs = synthetic.krotov_batch(B, np_, nt, seed=3)
gb = control.GridBatch(pbs, ctx=ctx)
pl = gb.plan
sqrt_hf_T = np.array([np.complex128(np.sqrt(p.laser_field_hf(np.float64(1.0), p.T, p.pcos, p.w_list))) for p in pbs])
E_presc = gb._pad([gb._prescribed_field(p, control.FORWARD) for p in pbs])
ref = {}
for var in variants:
    os.environ["QC_GRID_VAR"] = str(var)
    # (a) prescribed field
    pl.set_E_base(E_presc)
    ts = []
    for rep in range(4):
        pl.snapshot_load(0)
        ctx.synchronize(); ctx.timer_start()
        pl.sweep(0, engine.FIELD_PRESCRIBED, 1, nt, True, -1, -1)
        ts.append(ctx.timer_stop())
    ms_a = min(ts[1:])
    st_a = pl.get_state()
    # (b) Krotov: iteration 0 fills the trajectories, then one timed iteration
    pl.snapshot_load(0); pl.record_state(0, 0, 0)
    pl.sweep(0, engine.FIELD_PRESCRIBED, 1, nt, True, -1, 0)
    pl.goal_overlap(1, fetch=False); pl.krotov_terminal(1, sqrt_hf_T, 1)
    pl.sweep(1, engine.FIELD_KROTOV_BWD, 1, nt, True, 0, 1)
    tf, tb = [], []
    for rep in range(3):
        pl.snapshot_load(0); pl.record_state(0, 0, 0)
        ctx.synchronize(); ctx.timer_start()
        pl.sweep(0, engine.FIELD_KROTOV_FWD, 0, nt + 1, True, 1, 0)
        tf.append(ctx.timer_stop())
        pl.goal_overlap(1, fetch=False); pl.krotov_terminal(1, sqrt_hf_T, 1)
        ctx.synchronize(); ctx.timer_start()
        pl.sweep(1, engine.FIELD_KROTOV_BWD, 1, nt, True, 0, 1)
        tb.append(ctx.timer_stop())
    st_b = pl.get_state()
    E_b = pl.get_fields()
    line = {"np": np_, "runs": B, "nt": nt, "QC_GRID_VAR": var, "prescribed_ms": ms_a,
            "prescribed_frac_fp64_peak": F_presc / (ms_a * 1e-3) / 1e12 / peak,
            "krotov_fwd_ms": min(tf[1:]), "krotov_fwd_frac_fp64_peak": F_krot / (min(tf[1:]) * 1e-3) / 1e12 / peak,
            "krotov_bwd_ms": min(tb[1:]), "fp64_peak_tflops_measured": peak}
    if not ref:
        ref = {"a": st_a, "b": st_b, "E": E_b}
    else:
        dx = pbs[0].dx
        line["max_l2_vs_var0_prescribed"] = float(np.sqrt(np.max(np.sum(np.abs(st_a - ref["a"]) ** 2, axis=(1, 2, 3)) * dx)))
        line["max_l2_vs_var0_krotov"] = float(np.sqrt(np.max(np.sum(np.abs(st_b - ref["b"]) ** 2, axis=(1, 2, 3)) * dx)))
        line["max_rel_field_vs_var0"] = float(np.max(np.abs(E_b - ref["E"])) / np.max(np.abs(ref["E"])))
    print(json.dumps(line), flush=True)
gb.close()
ctx.close()
