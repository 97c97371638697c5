"""Kernel B on the reference's 400-run T-sweep: one backward + one forward UT sweep timed with CUDA events, with the
warp retirement on / off (QC_NLEVEL_RETIRE) and the runs ordered by length or in the template's order."""
import contextlib, io, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from qcontrol_b200 import batch, config, control, engine, newcheb, task_manager

ctx = engine.Context(0)
variants = batch.expand_substitutions(bench.ut_jx_template(4))
pbs = []
for i, v in enumerate(variants):
    with contextlib.redirect_stdout(io.StringIO()):
        c = config.TaskRootConfiguration(); c.load(v)
        batch.apply_auto_rules(c, i)
        tm = task_manager.create(c)
    pbs.append(newcheb.problem_from(c, tm))
for order in ("template", "by_length"):
    runs = pbs if order == "template" else sorted(pbs, key=lambda p: -p.nt)
    for retire in ("1", "0", "1", "0"):
        os.environ["QC_NLEVEL_RETIRE"] = retire
        nb = control.NLevelBatch(runs, ctx=ctx)
        nb.time_sweeps, nb.sweep_ms = True, 0.0
        nb.run_unit_transform(iter_max=3, keep_fields=False)
        print(json.dumps({"order": order, "QC_NLEVEL_RETIRE": retire, "sweep_ms_per_iteration": nb.sweep_ms / 4}), flush=True)
        nb.close()
