"""Write the results of a few fixed workloads to an .npz (argv[1]) -- used to compare two builds of the library
bit for bit (QCONTROL_B200_LIB selects the build):  grid sweeps with renormalisation and rotating frame at every
size, and the unitary-transformation loop of the 2-level Jx test configuration."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcontrol_b200 import control, engine, synthetic
from oracle import qc_oracle as qo
from tests.golden.gen_golden import conf_ut_jx

ctx = engine.Context(0)
out = {}
for np_ in (256, 512, 1024, 2048):
    for B in (6, 600 if np_ <= 512 else 6):
        pbs = synthetic.krotov_batch(B, np_, 12, seed=5)
        gb = control.GridBatch(pbs, ctx=ctx, store_traj=False)
        pl = gb.plan
        pl.set_E_base(gb._pad([gb._prescribed_field(p, control.FORWARD) for p in pbs]))
        pl.snapshot_load(0)
        pl.sweep(0, engine.FIELD_PRESCRIBED, 1, 12, True, -1, -1)
        out["grid_%d_%d" % (np_, B)] = pl.get_state()
        gb.close()
pb = qo.make_problem(conf_ut_jx(400, 3))
nb = control.NLevelBatch([pb, pb], ctx=ctx)
hist = nb.run_unit_transform(iter_max=3)
out["ut_E"] = np.stack([r.E_tlist for r in hist[0]])
out["ut_tab"] = np.array([[r.it, r.goal_close, r.Fsm, r.E_int, r.J] for r in hist[0]])
nb.close()
np.savez(sys.argv[1], **out)
print("wrote", sys.argv[1], sorted(out))
