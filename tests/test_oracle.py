"""Pin the CPU oracle (oracle/qc_oracle.py) against the reference.

Sources of truth, in order:
  1. literal known-answer vectors of the reference's unit tests
     (tests/math_base_tests.py:6-51 in the reference tree);
  2. fixtures produced by running the unmodified reference here
     (tests/golden/gen_golden.py -> *.npz);
  3. rows of the reference's shipped golden tables (test_data/*.py ->
     shipped_tables.npz / shipped_grid_records.npz).
All CPU-only; the whole file runs in well under two minutes.
"""
import ast
import math

import numpy as np
import pytest

from oracle import qc_oracle as qo


def relerr(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


# ---------------------------------------------------------------- unit KATs
def test_cprod_kat():
    # reference tests/math_base_tests.py:6-16
    s = np.array([1, 2, 3, 4])
    assert qo.cprod(s, s, 0.2) == 6
    ones = np.array([1, 1, 1, 1])
    assert np.vdot(ones, ones * s) * 0.2 == 2


def test_initak_kat(golden):
    # reference tests/math_base_tests.py:19-23
    dk = 10.0 * math.pi / 3.0
    dk2 = np.complex128(dk * dk + 0j)
    assert np.allclose(qo.initak(4, 0.2, 2, 1), np.array([0, -dk2, -4.0 * dk2, -dk2]))
    g = golden("kat_math_base")
    assert np.array_equal(qo.initak(4, 0.2, 2, 1), g["initak_4"])
    assert np.array_equal(qo.initak(16, 0.3, 1, 1), g["initak_16_o1"])
    assert np.array_equal(qo.initak(8, 0.3, 2, -2), g["initak_8_triv"])


def test_reorder_kat(golden):
    # reference tests/math_base_tests.py:32-39
    assert qo.reorder(2) == [0, 1]
    assert qo.reorder(8) == [0, 4, 2, 6, 1, 5, 3, 7]
    g = golden("kat_math_base")
    for nch in (2, 4, 8, 64):
        assert qo.reorder(nch) == list(g["reorder_%d" % nch])


def test_points_kat(golden):
    # reference tests/math_base_tests.py:42-51
    xp, dv = qo.points(4, 1)
    assert np.allclose(xp, [1.8477590650225735, -0.7653668647301795, 0.7653668647301797, -1.8477590650225735])
    assert np.allclose(dv, [-0.2734354014243603 - 0.9618903686220686j, -0.38060302857900286 - 0.6332232027087304j,
                            -0.3516313481159491 + 0.25126355493174846j, -2.1243197887715325e-17 + 0.135982856037932j])
    g = golden("kat_math_base")
    # bit-exact against the reference for the coefficient sets of every BASELINE config
    for nch, t in ((4, 1.0), (8, 5.65e-3), (8, -5.65e-3), (64, 0.2040), (64, 3.737), (64, -3.737), (64, 20.55)):
        tag = "points_%d_%s" % (nch, repr(t).replace("-", "m").replace(".", "p"))
        xp, dv = qo.points(nch, t)
        assert np.array_equal(xp, g[tag + "_xp"])
        assert np.array_equal(dv, g[tag + "_dv"])


# ---------------------------------------------------------------- one step
def _grid_problem(g):
    v = g["v"]
    return qo.Problem(hamil="ntriv", ntriv=1, np_=v.shape[1], nlevs=2, nb=1, dx=float(g["dx"]), v=v,
                      vmin=np.zeros(2), akx2=g["akx2"], nch=int(g["nch"]))


@pytest.mark.parametrize("name", ["harm", "morse", "krotov"])
def test_one_step_bitexact(golden, name):
    """prop_cpu / Hamil2DNonTrivial on the same input: the oracle performs the
    same numpy operations in the same order, so it must agree to the last bit."""
    g = golden("one_step_" + name)
    pb = _grid_problem(g)
    for tag, sign in (("f", 1.0), ("b", -1.0)):
        E = float(g["E_" + tag])
        out = qo.newton_propagate(pb, g["psi_in"], sign * float(g["t_sc"]), float(g["emin"]), float(g["emax"]),
                                  np.float64(E), float(g["eL"]), np.float64(E))
        assert np.array_equal(out, g["psi_out_" + tag])
    eL = float(g["eL"])
    assert np.array_equal(qo.apply_hamil(pb, g["psi_in"], np.float64(3.3), eL, np.complex128(3.3), False), g["hpsi_shift"])
    assert np.array_equal(qo.apply_hamil(pb, g["psi_in"], np.float64(3.3), eL, np.complex128(1.2 - 0.7j), True), g["hpsi_orig"])


# ---------------------------------------------------------------- control loops
def _problem_from(g, **over):
    user = ast.literal_eval(str(g["conf_json"]))
    user.update(over)
    return qo.make_problem(user), user


def _check_setup(pb, g):
    assert np.array_equal(pb.psi0, g["psi0"])
    assert np.array_equal(pb.psif, g["psif"])
    assert np.array_equal(pb.v, g["v"])
    assert np.array_equal(pb.vmin, g["vmin"])
    assert np.array_equal(pb.akx2, g["akx2"])
    assert np.array_equal(pb.x, g["x"])
    assert float(pb.dx) == float(g["dx"])
    assert float(pb.t_step) == float(g["t_step"])


def _run_and_compare(pb, g, keep=(), tol=0.0):
    recs = {}

    def on_step(pid, v, st):
        key = "%s__basis_%d" % (pid, v)
        if key + "__l" in g and st.rec.l in set(g[key + "__l"].tolist()):
            recs.setdefault(key, []).append(st.rec)

    fit = qo.Fitter(pb, on_step=on_step)
    err = ""
    try:
        rows = fit.run()
    except ValueError as e:
        err = str(e)
        rows = fit.rows
    assert err == str(g["error"])
    tab = np.array([[r.it, r.goal_close, r.Fsm, r.E_int, r.J] for r in rows])
    assert tab.shape == g["iter_tab"].shape
    if tol == 0.0:
        assert np.array_equal(tab, g["iter_tab"])
        assert np.array_equal(np.stack([r.E_tlist for r in rows]), g["iter_E"])
    else:
        assert relerr(tab, g["iter_tab"]) < tol
        assert relerr(np.stack([r.E_tlist for r in rows]), g["iter_E"]) < tol
    for key, rr in recs.items():
        ls = g[key + "__l"].tolist()
        for r in rr:
            k = ls.index(r.l)
            for name, val in (("psi", r.psi), ("ener", r.cener), ("norm", r.cnorm), ("overlp0", r.overlp0),
                              ("overlpf", r.overlpf)):
                if name == "psi" and key + "__psi_l" in g:        # wavefunctions kept only at a subset of steps
                    pl = g[key + "__psi_l"].tolist()
                    if r.l not in pl:
                        continue
                    ref = g[key + "__psi"][pl.index(r.l)]
                else:
                    ref = g[key + "__" + name][k]
                if tol == 0.0:
                    assert np.array_equal(val, ref), (key, r.l, name)
                else:
                    assert relerr(val, ref) < tol, (key, r.l, name)
            # the reporter receives |E| on grids and Re E otherwise (propagation.py:291-294)
            shown = abs(r.E) if pb.ntriv == 1 else np.real(r.E)
            assert complex(shown) == complex(g[key + "__E"][k])
            assert float(r.t) == float(g[key + "__t"][k])
            if key + "__freq_mult" in g:
                assert float(r.freq_mult) == float(g[key + "__freq_mult"][k]), (key, r.l)
    return fit


@pytest.mark.parametrize("name", ["fit_ut_jx", "fit_ut_jx_4lvls", "fit_ut_jz", "fit_ut_s2s", "fit_ut_jx_4lvls_wc",
                                  "fit_ut_jz_wc", "fit_ut_s2s_wc", "fit_ut_jx_re", "fit_ut_jx_ss",
                                  "fit_ut_jx_maxwell_cos", "fit_ut_jx_const_sin"])
def test_unit_transform_bitexact(golden, name):
    """Whole optimisation loops on N-level systems: backward/forward sweeps,
    field update from the stored trajectory, dynamic h_lambda, F / E_int / J."""
    g = golden(name)
    pb, _ = _problem_from(g)
    _check_setup(pb, g)
    _run_and_compare(pb, g)


def test_task_functions_kat(golden):
    """Every envelope, carrier, functional and 'a' coefficient of the oracle against values computed by the
    reference's own functions (task_manager.py:18-268) on the inputs of gen_golden.kat_task_inputs: bit-exact."""
    from tests.golden.gen_golden import kat_task_inputs
    g = golden("kat_task_functions")
    env_in, hf_in, gcs, bases = kat_task_inputs()
    for name in ("zero", "const", "gauss", "sqrsin", "maxwell"):
        got = np.array([qo.ENVELOPES[name](*a) for a in env_in], np.float64)
        assert np.array_equal(got, g["env_" + name]), name
    for name in ("exp", "cos", "sin", "cos_set", "sin_set"):
        got = np.array([qo.CARRIERS[name](*a) for a in hf_in], np.complex128)
        assert np.array_equal(got, g["hf_" + name]), name
    for kind in ("sm", "re", "ss"):
        F_fn, a_fn = qo.FUNCTIONALS[kind]
        for k, gc in enumerate(gcs):
            assert np.complex128(F_fn(gc, len(gc))) == g["F_%s_%d" % (kind, k)], (kind, k)
        for k, (psi, chi, dx) in enumerate(bases):
            assert np.array_equal(a_fn(psi, chi, dx), g["a_%s_%d" % (kind, k)]), (kind, k)


def test_pi_pulse_bitexact(golden):
    g = golden("fit_pi_pulse")
    pb, _ = _problem_from(g, _hamil_override="qubit", _ntriv_override=-2)
    pb.psif = g["psif"].copy()          # reference tests/functional_tests.py:45-50
    pb.init_dir = qo.FORWARD
    _check_setup(pb, g)
    _run_and_compare(pb, g)


@pytest.mark.parametrize("name", ["fit_single_harm", "fit_single_morse", "fit_trans_woc"])
def test_prescribed_field_grid_bitexact(golden, name):
    g = golden(name)
    pb, _ = _problem_from(g)
    _check_setup(pb, g)
    _run_and_compare(pb, g)


def test_krotov_256_bitexact(golden):
    g = golden("fit_krotov_256")
    pb, _ = _problem_from(g)
    _check_setup(pb, g)
    _run_and_compare(pb, g)


@pytest.mark.parametrize("name", ["fit_loc_pop_256", "fit_loc_proj_256"])
def test_local_control_bitexact(golden, name):
    """Local control: the Euler ODE of the field (population) and the fed-back carrier-frequency
    multiplier with its per-step energy window and Newton table (projection), every step."""
    g = golden(name)
    pb, _ = _problem_from(g)
    _check_setup(pb, g)
    fit = _run_and_compare(pb, g)
    if "proj" in name:
        fm = g["iter_0f__basis_0__freq_mult"]
        assert fm.min() == 1.0 and fm.max() > 1.9          # the feedback branch really fired


# ---------------------------------------------------------------- shipped tables
def test_shipped_ut_jx_table(golden):
    """The reference's own golden for test_opt_ctrl_ut_HB_2lvls_Jx
    (tests/functional_tests.py:879-972; test_data/fit_iter_opt_ctrl_ut_HB_2lvls_Jx.py),
    full length (nt = 3500, 4 iterations), with the reference's tolerances."""
    s = golden("shipped_tables")
    from tests.golden.gen_golden import conf_ut_jx
    pb = qo.make_problem(conf_ut_jx(3500, 3))
    rows = qo.Fitter(pb).run()
    tab = np.array([[r.it, r.goal_close, r.Fsm, r.E_int, r.J] for r in rows])
    ref = s["opt_ctrl_ut_HB_2lvls_Jx__iter_tab"]
    assert tab.shape == ref.shape
    assert np.array_equal(tab[:, 0], ref[:, 0])
    # iter_fit_comparer: goal_close 1e-4, Fsm 1e-5, E_int 1e-4, J 1e-5 (relative); we are far inside
    assert np.allclose(tab[1:, 1:], ref[1:, 1:], rtol=1e-9, atol=0)
    assert abs(tab[0, 1] - ref[0, 1]) < 1e-12 and np.allclose(tab[0, 2:], ref[0, 2:], rtol=1e-9, atol=1e-12)
    E = np.stack([r.E_tlist for r in rows])
    assert np.allclose(E, s["opt_ctrl_ut_HB_2lvls_Jx__iter_E"], rtol=1e-8, atol=1e-8)
    assert np.allclose(np.array(pb.t_list), s["opt_ctrl_ut_HB_2lvls_Jx__iter_t"], rtol=1e-15)


def test_shipped_pi_pulse_table(golden):
    s = golden("shipped_tables")
    g = golden("fit_pi_pulse")
    pb, _ = _problem_from(g, _hamil_override="qubit", _ntriv_override=-2)
    pb.psif = g["psif"].copy()
    pb.init_dir = qo.FORWARD
    rows = qo.Fitter(pb).run()
    tab = np.array([[r.it, r.goal_close, r.Fsm, r.E_int, r.J] for r in rows])
    ref = s["pi_pulse__iter_tab"]
    assert tab.shape == ref.shape
    assert np.allclose(tab, ref, rtol=1e-9, atol=1e-12)


def test_grid2d_separable():
    """The 2-D restatement (oracle/qc_oracle2d.py, no reference counterpart) against the pinned 1-D path: for
    V(x,y) = V1(x) + V2(y) and a product state, converged 2-D steps equal the outer product of 1-D steps."""
    from oracle import qc_oracle2d as q2
    N, L = 64, 10.0
    dx = L / (N - 1)
    x = (np.arange(N) - (N - 1) / 2) * dx
    ak = qo.initak(N, dx, 2, 1) * (-qo.HART_TO_CM / (2.0 * 0.5 * qo.DALT_TO_AU))
    v1, v2 = 9000 * x ** 2 / 2, 5000 * (x - 0.3) ** 2 / 2
    f = np.exp(-(x - 0.5) ** 2 / 0.8).astype(complex)
    g = (np.exp(-(x + 0.4) ** 2 / 1.1) * np.exp(2j * x)).astype(complex)
    f /= np.sqrt(dx * np.sum(abs(f) ** 2))
    g /= np.sqrt(dx * np.sum(abs(g) ** 2))
    v = v1[:, None] + v2[None, :]
    dt, nch, steps = 4e-18, 64, 5
    emax, emin = q2.energy_window_2d(v, ak, ak, N)
    psi = f[:, None] * g[None, :]
    for _ in range(steps):
        psi, nrm = q2.step_2d(psi, v, ak, ak, nch, dt, dx * dx, emin, emax)
        assert abs(nrm - 1.0) < 1e-10

    def p1(state, vv):
        pb = qo.Problem(hamil="ntriv", ntriv=1, np_=N, nlevs=2, nb=1, dx=dx, v=np.stack([vv, vv]), vmin=np.zeros(2),
                        akx2=ak, nch=nch)
        em = float(vv.max() + abs(ak[N // 2 - 1]))
        t_sc = dt * em * qo.CM_TO_ERG / 4 / qo.RED_PLANCK_H
        st = np.stack([state, np.zeros_like(state)])
        for _ in range(steps):
            st = qo.newton_propagate(pb, st, t_sc, 0.0, em, np.float64(0.0), 0.0, np.float64(0.0))
            st = st / np.sqrt(abs(np.vdot(st, st) * dx))
        return st[0]

    ref = p1(f, v1)[:, None] * p1(g, v2)[None, :]
    assert np.sqrt(dx * dx * np.sum(abs(psi - ref) ** 2)) < 1e-12
