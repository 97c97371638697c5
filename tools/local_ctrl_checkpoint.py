"""Run the shipped local_control_projection configuration (tests/functional_tests.py:599-635) on the GPU up to
step L0, store the state and the feedback variables there, continue to step L1 and store the multipliers of
every step in between.  The CPU oracle can then be started from the checkpoint (conditioning of
the projection feedback) without propagating the first L0 steps on the host.

    python tools/local_ctrl_checkpoint.py --l0 230000 --l1 240000 --out gpurun_out/loc_proj_ckpt.npz
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--l0", type=int, default=230000)
    ap.add_argument("--l1", type=int, default=240000)
    ap.add_argument("--out", default="gpurun_out/loc_proj_ckpt.npz")
    a = ap.parse_args()
    from oracle import qc_oracle as qo                       # problem set-up only (test infrastructure)
    from qcontrol_b200 import control, engine, math_base
    import tests.golden.gen_golden as gg
    pb = qo.make_problem(gg.conf_local_shipped("local_control_projection"))
    gb = control.GridBatch([pb], store_traj=False)
    pl = gb.plan
    kmax = abs(pb.akx2[int(pb.np_ / 2 - 1)])
    params = np.array([[pb.v[0][0] + kmax, pb.vmin[0], pb.v[1][0] + kmax, pb.vmin[1], pb.E0, pb.nu_L, pb.nu, pb.k_E,
                        pb.lamb, pb.pow_, pb.epsilon, pb.t_step]])
    pl.set_local_control(params, gb._pad([control.step_times(pb, control.FORWARD)], np.float64),
                         math_base.newton_table(pb.nch, 1.0)[0])
    pl.snapshot_load(0)
    pl.set_E_base(gb._pad([gb._prescribed_field(pb, control.FORWARD)]))
    pl.sweep(0, engine.FIELD_LOCAL_PROJ, 1, a.l0, True, -1, -1)
    psi0 = pl.get_state()[0, 0]
    st0 = pl.get_local_state()[0]
    pl.sweep(0, engine.FIELD_LOCAL_PROJ, a.l0 + 1, a.l1 - a.l0, True, -1, -1)
    psi1 = pl.get_state()[0, 0]
    st1 = pl.get_local_state()[0]
    fm = pl.get_fields()[0].imag[a.l0:a.l1 + 1]
    np.savez(a.out, l0=a.l0, l1=a.l1, psi0=psi0, state0=st0, psi1=psi1, state1=st1, freq_mult=fm)
    print("checkpoint", a.l0, st0[:4], "->", a.l1, st1[:4])


if __name__ == "__main__":
    main()
