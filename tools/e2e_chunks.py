"""e2e path (qc_propagate_host) with different pipeline chunk sizes; prints ms per call and grid-pt*timesteps/s."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qcontrol_b200 import control, engine, synthetic

B, nt = 8288, 16
pbs = synthetic.krotov_batch(B, 2048, nt, seed=12345)
ctx = engine.Context(0)
gb = control.GridBatch(pbs, ctx=ctx, store_traj=False)
pl = gb.plan
hin = torch.empty(pl.state_shape, dtype=torch.complex128).pin_memory()
hout = torch.empty(pl.state_shape, dtype=torch.complex128).pin_memory()
hin.numpy()[...] = gb.psi0
pl.set_E_base(gb._pad([gb._prescribed_field(p, control.FORWARD) for p in pbs]))
# device-only reference: the same sweep with the state already resident
pl.snapshot_load(0)
ctx.synchronize(); ctx.timer_start(); pl.sweep(0, engine.FIELD_PRESCRIBED, 1, nt, True, -1, -1); dev_ms = ctx.timer_stop()
print("device-only sweep %.1f ms" % dev_ms)
for chunk in (592, 148, 296, 1184, 2072, 4144, 100000):
    os.environ["QC_E2E_CHUNK"] = str(chunk)
    ts = []
    for i in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        pl.propagate_host(hin.numpy(), hout.numpy(), slot=0, l_first=1, nsteps=nt, renorm=True)
        ts.append(time.perf_counter() - t0)
    print("chunk %6d: %.1f ms  %.3e" % (chunk, min(ts[1:]) * 1e3, B * 2048 * nt / min(ts[1:])))
