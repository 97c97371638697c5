"""Microseconds per time step of ONE run of each shipped grid configuration (BASELINE configs 1-3 are single runs of
2-8e5 steps): the plain path (all nch = 64 terms, like the reference) and the opt-in truncated series
(QC_POLY_TOL-style bound of 1e-11 per step), one CTA per run.

    python tools/shipped_latency.py [steps]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import qc_oracle as qo            # problem set-up only (tool, not product)
from qcontrol_b200 import control, engine
from tests.golden.gen_golden import conf_krotov, conf_single_harm, conf_single_morse

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ctx = engine.Context(0)
for name, user in (("single_harm np=512", conf_single_harm(steps)), ("single_morse np=2048", conf_single_morse(steps)),
                   ("krotov np=1024 (prescribed sweep)", conf_krotov(steps, 1))):
    pb = qo.make_problem(user)
    row = {"config": name, "steps": steps}
    for tol in (None, 1e-11):
        gb = control.GridBatch([pb], ctx=ctx, store_traj=False, poly_tol=tol)
        pl = gb.plan
        pl.set_E_base(gb._pad([gb._prescribed_field(pb, control.FORWARD)]))
        ts = []
        for rep in range(3):
            pl.snapshot_load(0)
            ctx.synchronize(); ctx.timer_start()
            pl.sweep(0, engine.FIELD_PRESCRIBED, 1, steps, True, -1, -1)
            ts.append(ctx.timer_stop())
        key = "full_series" if tol is None else "truncated_1e-11"
        row[key + "_us_per_step"] = min(ts[1:]) * 1e3 / steps
        if tol:
            row["terms_of_64"] = int(gb.poly_orders[control.FORWARD][0])
            row["l2_distance_after_%d_steps" % steps] = float(np.sqrt(pb.dx * np.sum(np.abs(pl.get_state() - full) ** 2)))
        else:
            full = pl.get_state()
        gb.close()
    print(json.dumps(row), flush=True)
ctx.close()
