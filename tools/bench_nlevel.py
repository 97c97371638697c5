#!/usr/bin/env python
"""Throughput of kernel B (N-level unitary-transformation control) on the reference's batch family
(input_batch/inp_Jx_250-19000fs_25runs.json: 25 run ids x 16 durations = 400 runs, np = 1, nch = 8,
nb = nlevs = 2) and on a synthetic enlargement of it.  Prints one JSON line per batch size.

    python tools/bench_nlevel.py [--iters 3] [--copies 1 20]
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np

T_LIST = [2.5E-13, 11.8E-12, 13.4E-12, 14.0E-12, 15.0E-12, 16.6E-12, 16.8E-12, 17.0E-12, 17.6E-12, 17.8E-12, 18.0E-12,
          18.2E-12, 18.4E-12, 18.6E-12, 18.8E-12, 19.0E-12]


def template(iters):
    return {"run_id": {"@SUBST:LIST": ["ut/run%d" % i for i in range(1, 26)]},
            "task_type": "optimal_control_unit_transform", "pot_type": "none", "wf_type": "const", "hamil_type": "BH_model",
            "lf_aug_type": "x", "init_guess": "sqrsin", "init_guess_hf": "sin_set", "nb": 2, "T": {"@SUBST:LIST": T_LIST},
            "np": 1, "L": 1.0, "Du": 1.0, "U": 30.0, "W": 30.0, "delta": 15.0, "nt_auto": True, "sigma_auto": True,
            "nu_L_auto": True,
            "fitter": {"epsilon": 1e-30, "impulses_number": 1, "iter_max": iters - 1, "iter_mid_1": 250, "iter_mid_2": 300,
                       "q": 0.75, "h_lambda": 0.00005, "h_lambda_mode": "dynamical", "pcos": 2.0, "hf_hide": False,
                       "propagation": {"nch": 8, "t0": 0.0, "E0": 40.0}, "mod_log": 500}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--copies", type=int, nargs="+", default=[1, 20])
    a = ap.parse_args()
    from qcontrol_b200 import batch, config, control, engine, newcheb, task_manager
    ctx = engine.Context(0)
    variants = batch.expand_substitutions(template(a.iters))
    pbs = []
    for i, v in enumerate(variants):
        with contextlib.redirect_stdout(io.StringIO()):
            c = config.TaskRootConfiguration()
            c.load(v)
            batch.apply_auto_rules(c, i)
            tm = task_manager.create(c)
        pbs.append(newcheb.problem_from(c, tm))
    for copies in a.copies:
        runs = sorted(pbs * copies, key=lambda p: -p.nt)      # runs of similar length share warps (they retire together)
        nb = control.NLevelBatch(runs, ctx=ctx)
        # one call; iteration 0 carries one-time work (field tables of the backward sweep, first-touch of the host
        # arrays), so the rate is taken over iterations 1 ... iters-1 from the loop's own time marks
        nb.time_sweeps, nb.sweep_ms = True, 0.0
        ctx.synchronize()
        t0 = time.perf_counter()
        hist = nb.run_unit_transform(iter_max=a.iters - 1, keep_fields=False)
        ctx.synchronize()
        total = time.perf_counter() - t0
        marks = nb.iter_marks
        assert len(marks) == a.iters + 1
        d_it = a.iters - 1
        d_wall = marks[-1] - marks[1]
        d_sweep = nb.sweep_ms * 1e-3 * d_it / a.iters
        walls = {2: (0.0, 0.0, total - (marks[-1] - marks[0]))}
        steps = sum(p.nt for p in runs) * 2 * d_it * runs[0].nb          # solver steps (per basis vector)
        traj_bytes = sum(p.nt for p in runs) * d_it * 2 * runs[0].nb * runs[0].nlevs * 16
        print(json.dumps({"workload": "ut_jx_Tsweep x%d" % copies, "runs": len(runs), "iterations_timed": d_it,
                          "setup_s": walls[2][2] - walls[2][0], "wall_s_per_iteration": d_wall / d_it,
                          "sweep_kernels_s_per_iteration": d_sweep / d_it,
                          "ut_iterations_per_s": len(runs) * d_it / d_wall,
                          "solver_steps_per_s": steps / d_wall, "traj_hbm_gbs_algorithmic": traj_bytes / d_wall / 1e9,
                          "nt_range": [min(p.nt for p in runs), max(p.nt for p in runs)], "launches": ctx.launches}))
        nb.close()
    ctx.close()


if __name__ == "__main__":
    main()
