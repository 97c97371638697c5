"""GPU parity of the two local-control task types against fixtures recorded from the
unmodified reference (tests/golden/gen_golden.py, case fitter_local):

* local_control_population -- the field follows an Euler ODE towards the envelope (fitter.py:673-692);
* local_control_projection -- the carrier-frequency multiplier is fed back from the overlaps of the previous
  step (fitter.py:598-623, 814-836), so eL, exp_L, the energy window and the Newton table change every step
  and are rebuilt inside the sweep kernel (qcontrol_b200/csrc/qc_local.cuh).

Tolerances: state 1e-10 per step (accumulated linearly), scalars 1e-9 relative (north_star)."""
import ast
import contextlib
import io
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import qc_oracle as qo

KEY = "iter_0f__basis_0__"


def l2(a, b, dx=1.0):
    return float(np.sqrt(dx * np.sum(np.abs(np.asarray(a) - np.asarray(b)) ** 2)))


@pytest.fixture(scope="module")
def ctx():
    from qcontrol_b200 import engine
    c = engine.Context(0)
    yield c
    c.close()


def _problem(g):
    return qo.make_problem(ast.literal_eval(str(g["conf_json"])))


def _psi_records(g):
    ls = g[KEY + "psi_l"] if KEY + "psi_l" in g else g[KEY + "l"]
    return [int(l) for l in ls], g[KEY + "psi"]


def test_local_population_vs_reference_golden(ctx, golden):
    from qcontrol_b200 import control
    g = golden("fit_loc_pop_256")
    pb = _problem(g)
    gb = control.GridBatch([pb, pb], ctx=ctx, store_traj=False)
    # the field list: every step's E as the reference's reporter saw it
    E = np.array(gb._local_population_field(pb))
    assert np.array_equal(E[1:], g[KEY + "E"][1:].real)
    rows = gb.run_local_population()
    tab = np.array([[r.it, r.goal_close, r.Fsm, r.E_int, r.J] for r in rows[0]])
    assert np.allclose(tab, g["iter_tab"], rtol=1e-9, atol=1e-12), tab - g["iter_tab"]
    assert np.allclose(rows[0][0].E_tlist, g["iter_E"][0], rtol=1e-12, atol=0)
    ls, psis = _psi_records(g)
    out = gb.plan.get_state()
    assert ls[-1] == pb.nt
    assert l2(out[0, 0], psis[-1], pb.dx) <= 1e-10 * pb.nt
    assert np.array_equal(out[0], out[1])
    gb.close()


@pytest.mark.parametrize("name", ["fit_loc_proj_256", "fit_loc_proj_1024"])
def test_local_projection_vs_reference_golden(ctx, golden, name):
    from qcontrol_b200 import control, engine
    g = golden(name)
    pb = _problem(g)
    gb = control.GridBatch([pb, pb, pb], ctx=ctx, store_traj=False)
    rows = gb.run_local_projection()
    assert gb.errors == ["", "", ""]
    ref_l = g[KEY + "l"].astype(int)
    ref_fm = g[KEY + "freq_mult"]
    fm = gb.freq_mult[0]
    assert ref_fm.max() > 1.9                                   # the feedback branch fired in the reference ...
    err = np.abs(fm[ref_l] - ref_fm) / ref_fm
    assert err.max() <= 1e-9, (int(ref_l[err.argmax()]), err.max())   # ... and on the device, step for step
    assert np.array_equal(gb.freq_mult[0], gb.freq_mult[2])
    tab = np.array([[r.it, r.goal_close, r.Fsm, r.E_int, r.J] for r in rows[0]])
    assert np.allclose(tab, g["iter_tab"], rtol=1e-9, atol=1e-12), tab - g["iter_tab"]
    assert np.allclose(rows[0][0].E_tlist, g["iter_E"][0], rtol=1e-12, atol=0)
    st = gb.local_state[0]
    assert st[0] == fm[pb.nt] and st[4] == 0.0
    # wavefunctions at the recorded steps: re-run in chunks that end on them
    ls, psis = _psi_records(g)
    pl = gb.plan
    pl.snapshot_load(0)
    params = np.zeros((3, 12))
    kmax = abs(pb.akx2[int(pb.np_ / 2 - 1)])
    params[:] = [pb.v[0][0] + kmax, pb.vmin[0], pb.v[1][0] + kmax, pb.vmin[1], pb.E0, pb.nu_L, pb.nu, pb.k_E, pb.lamb,
                 pb.pow_, pb.epsilon, pb.t_step]
    from qcontrol_b200 import math_base
    pl.set_local_control(params, gb._pad([control.step_times(pb, control.FORWARD)] * 3, np.float64),
                         math_base.newton_table(pb.nch, 1.0)[0])
    prev = 0
    for k, l in enumerate(ls):
        if l == 0:
            continue
        pl.sweep(0, engine.FIELD_LOCAL_PROJ, prev + 1, l - prev, True, -1, -1)
        prev = l
        out = pl.get_state()
        e = l2(out[0, 0], psis[k], pb.dx)
        assert e <= 1e-10 * l, (name, l, e)
        assert np.array_equal(out[0], out[1])
    # chunked and single-launch sweeps give the same multipliers (state carried between launches)
    assert np.array_equal(pl.get_fields()[0].imag[1:], fm[1:])
    gb.close()


def test_local_projection_error_like_reference(ctx, golden):
    """A multiplier driven below zero stops the run with the reference's RuntimeError (fitter.py:820-821)."""
    from qcontrol_b200 import control
    g = golden("fit_loc_proj_256")
    user = ast.literal_eval(str(g["conf_json"]))
    user["fitter"]["k_E"] = 4e35            # k_E * dt^2 ~ 0.8, lamb * dt ~ 0.03: under-damped, overshoots below zero
    pb = qo.make_problem(user)
    done = []
    fit = qo.Fitter(pb, on_step=lambda pid, v, st: done.append(st.rec.l))
    with pytest.raises(RuntimeError) as ei:
        fit.run()
    gb = control.GridBatch([pb], ctx=ctx, store_traj=False)
    gb.run_local_projection()
    assert gb.errors[0] == str(ei.value)
    assert int(gb.local_state[0, 10]) == done[-1] + 1           # raised at the same step
    gb.close()


@pytest.mark.parametrize("name", ["fit_loc_pop_256", "fit_loc_proj_256"])
def test_fitting_solver_drop_in_local(golden, name):
    """The reference's call sequence (tests/functional_tests.py:508-661) through the drop-in FittingSolver with a
    recording reporter: iteration row, field list, and the time points (state, energies, multiplier)."""
    from qcontrol_b200 import config, fitter, task_manager
    from tests.test_gpu_dropin import _Rec
    g = golden(name)
    user = ast.literal_eval(str(g["conf_json"]))
    with contextlib.redirect_stdout(io.StringIO()):
        conf = config.TaskRootConfiguration()
        conf.load(user)
        tm = task_manager.create(conf)
    rep = _Rec(50)
    fs = fitter.FittingSolver(conf.fitter, conf.task_type, conf.T, conf.np, conf.L, tm.init_dir, tm.ntriv, tm.psi0, tm.psif,
                              tm.v, tm.akx2, tm.F_goal, tm.laser_field, tm.laser_field_hf, tm.F_type, tm.aF_type,
                              tm.hamil2D, rep, None, None)
    with contextlib.redirect_stdout(io.StringIO()):
        fs.time_propagation(tm.dx, tm.x, tm.t_step, tm.t_list)
    tab = np.array([r[:5] for r in rep.iter_rows], float)
    assert np.allclose(tab, g["iter_tab"], rtol=1e-9, atol=1e-12), tab - g["iter_tab"]
    rows = rep.props[os.path.join("iter_0f", "basis_0")].rows
    ref_l = g[KEY + "l"].tolist()
    ls, psis = _psi_records(g)
    assert [r["l"] for r in rows] == [l for l in ls if l % 50 == 0]
    for r in rows:
        k = ref_l.index(r["l"])
        l = max(1, r["l"])
        assert l2(r["psi"], psis[ls.index(r["l"])], float(tm.dx)) <= 1e-10 * l, (name, r["l"])
        for f, rt in (("ener", 1e-9), ("norm", 1e-9), ("overlp0", 1e-8), ("overlpf", 1e-8), ("momx", 1e-8), ("momx2", 1e-8)):
            ref = g[KEY + f][k]
            scale = max(1e-12, float(np.abs(ref).max()))
            assert np.abs(np.asarray(r[f]) - ref).max() <= rt * scale + 1e-11, (name, r["l"], f)
        assert abs(complex(r["E"]) - complex(g[KEY + "E"][k])) <= 1e-9 * max(1.0, abs(complex(g[KEY + "E"][k])))
        assert abs(float(r["freq_mult"]) - float(g[KEY + "freq_mult"][k])) <= 1e-9 * float(g[KEY + "freq_mult"][k])


def _ref_close(a, b, eps, delta=1e-12):
    """TableComparer.compare_floats of the reference (tests/test_tools.py:24-33)."""
    if abs(b) < delta and abs(a) < delta:
        return True
    if abs(b) < delta < abs(a):
        return False
    return abs(a - b) / abs(b) <= abs(eps)


@pytest.mark.parametrize("task,fam", [("local_control_population", "loc_ctrl_pop"),
                                      ("local_control_projection", "loc_ctrl_proj")])
def test_shipped_local_control_full_length(golden, task, fam):
    """The reference's own golden tables test_data/fitter_loc_ctrl_{pop,proj}.py (280 000 / 315 000 steps, a
    record every 10 000) with the tolerances of its functional tests (tests/functional_tests.py:583-585):
    t 1e-6, E 1e-5, freq_mult 1e-4, total energy 1e-7, total overlaps 1e-3.

    Projection: the multiplier feedback sets in near step 230 000 and from there the REFERENCE amplifies a
    3e-16 perturbation of psi to 3-4 % within 6 000 steps (tests/golden/gen_local_onset.py), so its table is
    compared with those tolerances up to step 230 000 and loosely (multiplier within 10 %, energy 1e-4)
    afterwards; the onset itself is pinned against the oracle in test_projection_feedback_onset_vs_oracle."""
    from qcontrol_b200 import config, fitter, task_manager
    from tests.test_gpu_dropin import _Rec
    import tests.golden.gen_golden as gg
    ref = golden("shipped_tables")[fam + "__tvals"]
    user = gg.conf_local_shipped(task)
    with contextlib.redirect_stdout(io.StringIO()):
        conf = config.TaskRootConfiguration()
        conf.load(user)
        tm = task_manager.create(conf)
    rep = _Rec(10000)
    fs = fitter.FittingSolver(conf.fitter, conf.task_type, conf.T, conf.np, conf.L, tm.init_dir, tm.ntriv, tm.psi0, tm.psif,
                              tm.v, tm.akx2, tm.F_goal, tm.laser_field, tm.laser_field_hf, tm.F_type, tm.aF_type,
                              tm.hamil2D, rep, None, None)
    with contextlib.redirect_stdout(io.StringIO()):
        fs.time_propagation(tm.dx, tm.x, tm.t_step, tm.t_list)
    rows = rep.props[os.path.join("iter_0f", "basis_0")].rows
    assert len(rows) == ref.shape[0]
    for r, want in zip(rows, ref):
        strict = task == "local_control_population" or r["l"] <= 230000
        got = (r["t"], r["E"], r["freq_mult"], np.sum(r["ener"]).real)
        for k, eps in enumerate((1e-6, 1e-5, 1e-4 if strict else 0.1, 1e-7 if strict else 1e-4)):
            assert _ref_close(float(np.real(got[k])), float(want[k].real), eps), (fam, r["l"], k, got[k], want[k])
        if not strict:
            continue
        for k, val in ((4, np.sum(r["overlp0"])), (5, np.sum(r["overlpf"]))):
            for part in ("real", "imag"):
                a, b = float(getattr(complex(val), part)), float(getattr(complex(want[k]), part))
                assert _ref_close(a, b, 1e-3) or abs(a - b) <= 1e-3 * abs(complex(want[k])), (fam, r["l"], k, val, want[k])


def test_projection_feedback_onset_vs_oracle(ctx, golden):
    """The shipped projection run restarted at step 230 000 (state and feedback variables from the fixture):
    the multipliers of the next 1 500 steps, where the feedback law of fitter.py:608-621 is active, against the
    oracle started from the same checkpoint.  The oracle's own sensitivity to a 3e-16 perturbation over that
    stretch (stored beside it) bounds what any implementation can reproduce."""
    from qcontrol_b200 import control, engine, math_base
    import tests.golden.gen_golden as gg
    g = golden("loc_proj_onset")
    pb = qo.make_problem(gg.conf_local_shipped("local_control_projection"))
    gb = control.GridBatch([pb], ctx=ctx, store_traj=False)
    pl = gb.plan
    kmax = abs(pb.akx2[int(pb.np_ / 2 - 1)])
    params = np.array([[pb.v[0][0] + kmax, pb.vmin[0], pb.v[1][0] + kmax, pb.vmin[1], pb.E0, pb.nu_L, pb.nu, pb.k_E,
                        pb.lamb, pb.pow_, pb.epsilon, pb.t_step]])
    pl.set_local_control(params, gb._pad([control.step_times(pb, control.FORWARD)], np.float64),
                         math_base.newton_table(pb.nch, 1.0)[0])
    pl.set_local_state(g["state0"][None, :])
    pl.set_state(g["psi0"][None, None])
    pl.set_E_base(gb._pad([gb._prescribed_field(pb, control.FORWARD)]))
    l0, n = int(g["l0"]), 1500
    pl.sweep(0, engine.FIELD_LOCAL_PROJ, l0 + 1, n, True, -1, -1)
    fm = pl.get_fields()[0].imag[l0 + 1:l0 + 1 + n]
    ref, pert = g["fm_oracle"][:n], g["fm_oracle_perturbed"][:n]
    assert ref[-1] > 1.08                                    # the multiplier is moving
    err = np.abs(fm - ref) / ref
    own = np.abs(pert - ref) / ref
    assert err[:1000].max() <= 1e-12, err[:1000].max()
    assert err.max() <= 1e-9, err.max()                      # north_star tolerance over the whole stretch
    assert err.max() <= 10 * max(own.max(), 1e-12)           # and no worse than the reference's own conditioning
    gb.close()
