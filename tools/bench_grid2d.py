"""Kernel C (512 x 512 one-level grid): time per step, grid-pt*timesteps/s, FP64 and L2-traffic figures."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcontrol_b200 import engine, math_base
from qcontrol_b200.phys_consts import CM_TO_ERG, RED_PLANCK_H, HART_TO_CM, DALT_TO_AU

N, nch = 512, 64
ctx = engine.Context(0)
L = 10.0
dx = L / (N - 1)
x = (np.arange(N) - (N - 1) / 2) * dx
ak = np.ascontiguousarray((math_base.initak(N, dx, 2, 1) * (-HART_TO_CM / (2.0 * 0.5 * DALT_TO_AU))).real)
X, Y = np.meshgrid(x, x, indexing="ij")
v = 3000.0 * (X ** 2 + 1.1 * Y ** 2) + 800.0 * X * Y
emax, emin = float(v.max() + 2 * abs(ak[N // 2 - 1])), 0.0
dt = 1e-18
t_sc = dt * (emax - emin) * CM_TO_ERG / 4.0 / RED_PLANCK_H
xp, dv = math_base.newton_table(nch, t_sc)
peak = ctx.fp64_peak_tflops()
for batch in (1, 2, 8):
    psi = np.exp(-((X - 0.3) ** 2 + Y ** 2) / 2.0).astype(np.complex128)
    psi /= np.sqrt(dx * dx * np.sum(np.abs(psi) ** 2))
    pl = engine.Plan2D(ctx, N, nch, batch)
    pl.set_static(v, ak, ak, dx * dx)
    pl.set_poly(xp, dv, emin, emax, t_sc)
    pl.set_state(np.broadcast_to(psi, (batch, N, N)).copy())
    pl.sweep(2)
    steps = 20
    ctx.synchronize(); ctx.timer_start(); pl.sweep(steps); ms = ctx.timer_stop()
    per_step = ms / steps * 1e3
    flops = (nch - 1) * N * N * (2 * 10 * 9 + 22) + 18 * N * N          # two 1-D transform pairs per point per order
    print(json.dumps({"workload": "grid2d_512x512", "wavefunctions": batch, "us_per_step_per_launch_wave": per_step / -(-batch // 2),
                      "grid_pt_timesteps_per_s": batch * N * N * steps / (ms * 1e-3),
                      "tflops_algorithmic": batch * flops * steps / (ms * 1e-3) / 1e12, "fp64_peak_measured": peak,
                      "frac": batch * flops * steps / (ms * 1e-3) / 1e12 / peak,
                      "l2_bytes_per_order": 5 * N * N * 16 + N * N * 8, "norm": float(pl.get_norms()[0])}), flush=True)
    pl.close()
