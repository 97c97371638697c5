"""GPU tests of the reference-shaped layer: prop_cpu, Hamil2D functors, PropagationSolver,
FittingSolver and the batch entry point, driven the way the reference's own tests drive them
(tests/unit_tests_propagation.py, tests/functional_tests.py) and compared with fixtures recorded from
the unmodified reference."""
import ast
import contextlib
import copy
import io
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import qc_oracle as qo


def l2(a, b, dx=1.0):
    return float(np.sqrt(dx * np.sum(np.abs(np.asarray(a) - np.asarray(b)) ** 2)))


def _setup(user):
    from qcontrol_b200 import config, task_manager
    with contextlib.redirect_stdout(io.StringIO()):
        c = config.TaskRootConfiguration()
        c.load(user)
        return c, task_manager.create(c)


@pytest.mark.parametrize("name,maker,nt", [("harm", "conf_single_harm", 2000), ("morse", "conf_single_morse", 2000),
                                           ("krotov", "conf_krotov", 2000)])
def test_prop_cpu_and_hamil_functor_drop_in(golden, name, maker, nt):
    """S1: phys_base.prop_cpu(psi, hamil2D, t_sc, nch, np, emin, emax, E, eL, E_full, orig) and
    hamil2D(orig, psi, E, eL, E_full) with the reference's call signatures."""
    from qcontrol_b200 import phys_base
    from qcontrol_b200.psi_basis import Psi
    import tests.golden.gen_golden as gg
    g = golden("one_step_" + name)
    user = getattr(gg, maker)(nt) if name != "krotov" else gg.conf_krotov(nt, 1)
    conf, tm = _setup(user)
    dx = float(g["dx"])
    for tag, sign in (("f", 1.0), ("b", -1.0)):
        psi = Psi(f=[g["psi_in"][0].copy(), g["psi_in"][1].copy()], lvls=2)
        E = np.float64(g["E_" + tag])
        out = phys_base.prop_cpu(psi=psi, hamil2D=tm.hamil2D, t_sc=sign * float(g["t_sc"]), nch=int(g["nch"]), np=conf.np,
                                 emin=float(g["emin"]), emax=float(g["emax"]), E=E, eL=float(g["eL"]), E_full=E, orig=False)
        assert out is psi
        assert l2(out.as_array(), g["psi_out_" + tag], dx) <= 1e-10
    psi = Psi(f=[g["psi_in"][0].copy(), g["psi_in"][1].copy()], lvls=2)
    # H psi contains the kinetic term k^2/2m * psi_hat with k^2/2m up to emax: round-off of either FFT
    # is ~ eps * emax * |psi|, whatever the size of the (strongly cancelling) result
    tol = 2e-15 * float(g["emax"]) * float(np.abs(g["psi_in"]).max())
    h1 = tm.hamil2D(False, psi, np.float64(3.3), float(g["eL"]), np.complex128(3.3))
    assert np.abs(h1.as_array() - g["hpsi_shift"]).max() <= tol
    h2 = tm.hamil2D(True, psi, np.float64(3.3), float(g["eL"]), np.complex128(1.2 - 0.7j))
    assert np.abs(h2.as_array() - g["hpsi_orig"]).max() <= tol
    assert np.array_equal(psi.as_array(), g["psi_in"])          # the functor does not touch its argument


class _Rec:
    """Recording reporters, like tests/test_tools.py:124-286 of the reference."""

    def __init__(self, mod):
        self.mod = mod
        self.iter_rows, self.props = [], {}

    def open(self):
        return self

    def close(self):
        pass

    def print_iter_point_fitter(self, it, goal_close, E_tlist, t_list, Fsm, E_int, J, nt):
        self.iter_rows.append((it, goal_close, Fsm.real, E_int.real, J, np.array(E_tlist)))

    def create_propagation_reporter(self, prop_id, nlevs):
        outer = self

        class P:
            mod_fileout = outer.mod
            lmin = 0

            def __init__(self):
                self.rows = []

            def open(self):
                return self

            def close(self):
                pass

            def print_time_point_prop(self, l, psi, t, x, np_, nt, moms, smoms, ener, norm, overlp0, overlpf, overlp_tot,
                                      ener_tot, psi_max_abs, psi_max_real, E, freq_mult):
                if l % self.mod_fileout == 0:
                    self.rows.append(dict(l=l, t=t, psi=psi.as_array().copy(), momx=np.array(moms.x), momx2=np.array(moms.x2),
                                          momp=np.array(moms.p), momp2=np.array(moms.p2), ener=np.array(ener),
                                          norm=np.array(norm), overlp0=np.array(overlp0), overlpf=np.array(overlpf), E=E,
                                          psi_max_abs=np.array(psi_max_abs), psi_max_real=np.array(psi_max_real),
                                          freq_mult=freq_mult))
        p = P()
        self.props[prop_id] = p
        return p


def _compare_records(rows, g, key, dx, psi_tol):
    ls = g[key + "__l"].tolist()
    assert [r["l"] for r in rows] == ls
    for k, r in enumerate(rows):
        l = max(1, r["l"])
        assert l2(r["psi"], g[key + "__psi"][k], dx) <= psi_tol * l, (key, r["l"])
        for f, rt in (("ener", 1e-9), ("norm", 1e-9), ("overlp0", 1e-8), ("overlpf", 1e-8), ("momx", 1e-8), ("momx2", 1e-8),
                      ("momp2", 1e-8), ("psi_max_abs", 1e-9)):
            ref = g[key + "__" + f][k]
            scale = max(1e-12, float(np.abs(ref).max()))
            assert np.abs(np.asarray(r[f]) - ref).max() <= rt * scale + 1e-11, (key, r["l"], f, r[f], ref)
        ref = g[key + "__momp"][k]
        assert np.abs(np.asarray(r["momp"]) - ref).max() <= 1e-7 * max(1.0, float(np.abs(g[key + "__momp2"][k]).max()) ** 0.5)
        assert abs(complex(r["E"]) - complex(g[key + "__E"][k])) <= 1e-9 * max(1.0, abs(complex(g[key + "__E"][k])))
        assert float(r["t"]) == float(g[key + "__t"][k])


def test_propagation_solver_drop_in(golden):
    """S2: PropagationSolver(...).start(...); while solver.step(0.0) -- the protocol of the reference's
    tests/unit_tests_propagation.py:115-191, with the instrumentation (energies, overlaps, moments)
    compared as well."""
    from qcontrol_b200 import task_manager
    from qcontrol_b200.propagation import PropagationSolver
    g = golden("fit_trans_woc")
    user = ast.literal_eval(str(g["conf_json"]))
    conf, tm = _setup(user)
    cp = conf.fitter.propagation
    rec = _Rec(50).create_propagation_reporter("x", 2)

    def envelope(prop, stat, dyn):
        return task_manager._LaserFields.laser_field_gauss(cp.E0, dyn.t, cp.t0, cp.sigma)

    def factory(l, t, psi, psi_omega, E, fm, d):
        return PropagationSolver.DynamicState(l, t, copy.deepcopy(psi), copy.deepcopy(psi_omega), E, fm, d)

    s = PropagationSolver(T=conf.T, np=conf.np, L=conf.L, _warning_collocation_points=None, _warning_time_steps=None,
                          reporter=rec, hamil2D=tm.hamil2D, laser_field_envelope=envelope, laser_field_hf=tm.laser_field_hf,
                          freq_multiplier=lambda d, st: np.float64(1.0), dynamic_state_factory=factory,
                          pcos=conf.fitter.pcos, w_list=conf.fitter.w_list, mod_log=500, ntriv=tm.ntriv,
                          hf_hide=conf.fitter.hf_hide, conf_prop=cp)
    s.start(tm.v, tm.akx2, tm.dx, tm.x, tm.t_step, tm.psi0.psis[0], tm.psif.psis[0], PropagationSolver.Direction.FORWARD)
    n = 0
    while s.step(0.0):
        n += 1
    assert n == cp.nt - 1 and s.dyn.l == cp.nt + 1
    _compare_records(rec.rows, g, "iter_0f__basis_0", float(tm.dx), 1e-10)


@pytest.mark.parametrize("name,mod", [("fit_krotov_256", 40), ("fit_single_harm", 10), ("fit_ut_jx", 100),
                                      ("fit_ut_jx_re", 100), ("fit_ut_jx_ss", 100), ("fit_ut_jx_maxwell_cos", 100),
                                      ("fit_ut_jx_const_sin", 100)])
def test_fitting_solver_drop_in(golden, name, mod):
    """S3: FittingSolver(...).time_propagation(dx, x, t_step, t_list) with recording reporters -- the
    call sequence of tests/functional_tests.py:120-133 -- against the reference's iteration rows and
    the time points its reporters recorded."""
    from qcontrol_b200 import fitter
    g = golden(name)
    user = ast.literal_eval(str(g["conf_json"]))
    conf, tm = _setup(user)
    rep = _Rec(mod)
    fs = fitter.FittingSolver(conf.fitter, conf.task_type, conf.T, conf.np, conf.L, tm.init_dir, tm.ntriv, tm.psi0, tm.psif,
                              tm.v, tm.akx2, tm.F_goal, tm.laser_field, tm.laser_field_hf, tm.F_type, tm.aF_type,
                              tm.hamil2D, rep, None, None)
    with contextlib.redirect_stdout(io.StringIO()):
        fs.time_propagation(tm.dx, tm.x, tm.t_step, tm.t_list)
    tab = np.array([r[:5] for r in rep.iter_rows], float)
    ref = g["iter_tab"]
    assert tab.shape == ref.shape
    assert np.allclose(tab[:, 0], ref[:, 0])
    assert np.allclose(tab[:, 1:], ref[:, 1:], rtol=1e-9, atol=1e-12), tab - ref
    E = np.stack([r[5] for r in rep.iter_rows])
    scale = np.abs(g["iter_E"]).max(axis=1, keepdims=True)
    assert np.all(np.abs(E - g["iter_E"]) <= 1e-9 * scale + 1e-300)
    checked = 0
    for pid, p in rep.props.items():
        key = pid.replace(os.sep, "__")
        if key + "__l" in g:
            _compare_records(p.rows, g, key, float(tm.dx), 1e-10 if conf.np > 1 else 1e-12)
            checked += 1
    assert checked >= 1
    assert fs.dyn.iter_step == int(ref[-1, 0]) and abs(fs.dyn.Fsm.real - ref[-1, 2]) <= 1e-9 * abs(ref[-1, 2])


def test_batch_entry_point(tmp_path):
    """S4: the newcheb batch path on a small "@SUBST:LIST" template (2 run ids x 2 durations of the UT Jx
    family, nt_auto / sigma_auto / nu_L_auto rules, seeded w_list): tables written per variant in the
    reference's file format, values against the oracle run with the same pinned w_list."""
    from qcontrol_b200 import batch, newcheb
    from tests.golden.gen_golden import conf_ut_jx
    tpl = conf_ut_jx(1, 2)
    tpl["run_id"] = {"@SUBST:LIST": ["r1", "r2"]}
    tpl["T"] = {"@SUBST:LIST": [2.5e-13, 4.0e-13]}
    tpl["nt_auto"] = True
    tpl["fitter"]["w_list"] = []
    tpl["fitter"]["h_lambda"] = 5e-4          # well-conditioned (see DESIGN.md "Conditioning")
    rep = {"fitter": {"out_path": str(tmp_path / "out_{input:$.run_id}" / "T={input:$.T}"), "table.imin": -1,
                      "propagation": {"table.mod_fileout": 100}}}
    variants = batch.expand_substitutions(tpl)
    assert len(variants) == 4
    res = newcheb.run_variants(variants, rep, seed_base=1000, device=0, verbose=False)
    for i, var in enumerate(variants):
        user = copy.deepcopy(var)
        user["fitter"]["w_list"] = list(np.random.default_rng(1000 + i).uniform(-2.0, 2.0, 3))
        pb = qo.make_problem(user, apply_auto=True)
        rows = qo.Fitter(pb).run()
        tab = np.array([[r.it, r.goal_close, r.Fsm, r.E_int, r.J] for r in rows])
        got = np.array([[r.it, r.goal_close, r.Fsm, r.E_int, r.J] for r in res[i]])
        assert got.shape == tab.shape and np.allclose(got[1:], tab[1:], rtol=1e-8, atol=0), (i, got - tab)
        path = tmp_path / ("out_%s" % var["run_id"]) / ("T=%s" % var["T"]) / "tab_iter.csv"
        lines = path.read_text().strip().splitlines()
        assert len(lines) == len(rows)
        assert lines[1].split()[0] == "0" and abs(float(lines[1].split()[2]) - tab[1, 2]) < 1e-6
        nE = (tmp_path / ("out_%s" % var["run_id"]) / ("T=%s" % var["T"]) / "tab_iter_E.csv").read_text().count("\n")
        assert nE == len(rows) * (pb.nt + 1)
        # the parameter sheet post_processing.py parses, and the per-step tables of a sweep (plotting_flag default = all)
        sheet = (path.parent / "table_inp_-1.txt").read_text()
        assert "nb:\t\t\t2\n" in sheet and "iter_mid_2:\t\t300\n" in sheet and "epsilon:\t\t" in sheet
        fit_tab = (path.parent / "iter_1f" / "basis_1" / "tables" / "tab_tvals_fit.csv").read_text().strip().splitlines()
        assert len(fit_tab) == pb.nt // 100 + 1
        assert abs(float(fit_tab[1].split()[0]) - 100 * pb.t_step * 1e15) < 1e-6
        abs0 = (path.parent / "iter_0b" / "basis_0" / "tables" / "tab_abs_0.csv").read_text().strip().splitlines()
        assert len(abs0) == pb.nt // 100 + 1


def test_batch_entry_point_pipelined_pieces(tmp_path, monkeypatch):
    """newcheb.run_variants with the batch cut into pieces (set-up of piece k + 1 on a second thread and stream beside
    the control loops of piece k, plans cut from the arena of qc_pool_reserve): a template of 8 Krotov variants at
    np = 256 in pieces of 3 (3 + 3 + 2, QC_PIPELINE_PIECE) must give exactly the tables of the uncut batch -- a run does
    not depend on its position in a batch -- and those of the oracle."""
    from qcontrol_b200 import batch, newcheb
    dt = 350e-15 / 230000
    nt = 60
    T = dt * nt
    tpl = {"task_type": "optimal_control_krotov", "pot_type": "morse", "wf_type": "morse", "hamil_type": "ntriv",
           "init_guess": "gauss", "init_guess_hf": "exp", "nb": 1, "nlevs": 2, "T": T, "L": 5.0, "np": 256, "a": 1.0,
           "De": 20000.0, "x0p": -0.17, "a_e": 1.0, "De_e": 10000.0, "Du": 20000.0,
           "fitter": {"epsilon": 1e-8, "impulses_number": 1, "iter_max": 2, "h_lambda": 0.0066, "mod_log": 10 ** 9,
                      "propagation": {"m": 0.5, "nch": 32, "nt": nt,
                                      "E0": {"@SUBST:LIST": [2000.0, 2400.0, 2861.6, 3200.0]},
                                      "t0": {"@SUBST:LIST": [T * 0.45, T * 0.55]},
                                      "sigma": T / 6, "nu_L": 0.29297e15}}}
    variants = batch.expand_substitutions(tpl)
    assert len(variants) == 8
    out = {}
    for piece in ("0", "3"):
        monkeypatch.setenv("QC_PIPELINE_PIECE", piece)
        rep = {"fitter": {"out_path": str(tmp_path / ("p" + piece) / "E0={input:$.fitter.propagation.E0}_t0={input:$.fitter.propagation.t0}"),
                          "plotting_flag": "tables_iter", "table.imin": -1}}
        tm = {}
        res = newcheb.run_variants(variants, rep, device=0, verbose=False, timings=tm)
        assert tm["plans"] == (1 if piece == "0" else 3)
        out[piece] = {i: np.array([[r.it, r.goal_close, r.Fsm, r.E_int, r.J] for r in rows]) for i, rows in res.items()}
        assert sorted(out[piece]) == list(range(8))
    for i in range(8):
        assert out["0"][i].shape == out["3"][i].shape == (3, 5)
        assert np.array_equal(out["0"][i], out["3"][i]), (i, out["0"][i] - out["3"][i])
    for i in (0, 5, 7):
        pb = qo.make_problem(copy.deepcopy(variants[i]))
        rows = qo.Fitter(pb).run()
        tab = np.array([[r.it, r.goal_close, r.Fsm, r.E_int, r.J] for r in rows])
        assert np.allclose(out["3"][i], tab, rtol=1e-9, atol=1e-12), (i, out["3"][i] - tab)


def test_newcheb_command_line(tmp_path):
    """`python -m qcontrol_b200.newcheb --json_task T.json --json_rep R.json` (newcheb.py:612-690): the reference's
    command line on a small substitution template; one output directory per variant, the summary table on stdout."""
    import json
    import subprocess
    import sys
    from tests.golden.gen_golden import conf_ut_jx
    tpl = conf_ut_jx(1, 1)
    tpl["run_id"] = "cli"
    tpl["T"] = {"@SUBST:LIST": [2.0e-13, 3.0e-13, 2.6e-13]}
    tpl["nt_auto"] = True
    tpl["fitter"]["h_lambda"] = 5e-4
    rep = {"fitter": {"out_path": str(tmp_path / "o_{input:$.run_id}" / "T={input:$.T}"), "plotting_flag": "tables_iter",
                      "table.imin": -1, "propagation": {"table.mod_fileout": 100}}}
    (tmp_path / "task.json").write_text(json.dumps(tpl))
    (tmp_path / "rep.json").write_text(json.dumps(rep))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "qcontrol_b200.newcheb", "--json_task", str(tmp_path / "task.json"),
                          "--json_rep", str(tmp_path / "rep.json"), "--seed_base", "1000"], cwd=root, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.count("Job #") == 3 and "goal_close" in out.stdout
    for T in (2.0e-13, 3.0e-13, 2.6e-13):
        d = tmp_path / "o_cli" / ("T=%s" % T)
        rows = (d / "tab_iter.csv").read_text().strip().splitlines()
        assert [r.split()[0] for r in rows] == ["-1", "0", "1"]
        assert not (d / "tab_iter_E.csv").exists()             # tables_iter: no field table (reporter.py:604-606)
        assert (d / "table_inp_-1.txt").exists()
    bad = subprocess.run([sys.executable, "-m", "qcontrol_b200.newcheb", "--json_task", str(tmp_path / "task.json")],
                         cwd=root, capture_output=True, text=True, timeout=120)
    assert bad.returncode == 2 and "Usage" in bad.stdout


def test_two_fitting_solvers_from_a_thread_pool(golden):
    """The reference farms its jobs over ThreadPoolExecutor (newcheb.py:656-659): two FittingSolvers running at the
    same time from two host threads -- each thread gets its own context and stream (phys_base.default_context is
    thread-local) -- must both reproduce the reference's tables; the one-step drop-ins prop_cpu / hamil2D called
    concurrently on the SAME functor object must not mix their states."""
    from concurrent.futures import ThreadPoolExecutor
    from qcontrol_b200 import fitter, phys_base
    from qcontrol_b200.psi_basis import Psi

    def job(name):
        g = golden(name)
        conf, tm = _setup(ast.literal_eval(str(g["conf_json"])))
        rep = _Rec(100)
        fs = fitter.FittingSolver(conf.fitter, conf.task_type, conf.T, conf.np, conf.L, tm.init_dir, tm.ntriv, tm.psi0,
                                  tm.psif, tm.v, tm.akx2, tm.F_goal, tm.laser_field, tm.laser_field_hf, tm.F_type,
                                  tm.aF_type, tm.hamil2D, rep, None, None)
        with contextlib.redirect_stdout(io.StringIO()):
            fs.time_propagation(tm.dx, tm.x, tm.t_step, tm.t_list)
        return np.array([r[:5] for r in rep.iter_rows], float), id(phys_base.default_context())

    names = ["fit_krotov_256", "fit_ut_jx", "fit_krotov_256", "fit_ut_jx"]
    with ThreadPoolExecutor(2) as pool:
        results = list(pool.map(job, names))
    for name, (tab, _) in zip(names, results):
        ref = golden(name)["iter_tab"]
        assert tab.shape == ref.shape and np.allclose(tab[:, 1:], ref[:, 1:], rtol=1e-9, atol=1e-12), name
    assert len({cid for _, cid in results}) == 2          # one context per worker thread
    assert id(phys_base.default_context()) not in {cid for _, cid in results}

    g = golden("one_step_krotov")
    import tests.golden.gen_golden as gg
    conf, tm = _setup(gg.conf_krotov(2000, 1))

    def step(k):
        psi = Psi(f=[g["psi_in"][0].copy() * (1 + k), g["psi_in"][1].copy() * (1 + k)], lvls=2)
        out = phys_base.prop_cpu(psi=psi, hamil2D=tm.hamil2D, t_sc=float(g["t_sc"]), nch=int(g["nch"]), np=conf.np,
                                 emin=float(g["emin"]), emax=float(g["emax"]), E=np.float64(g["E_f"]), eL=float(g["eL"]),
                                 E_full=np.float64(g["E_f"]), orig=False)
        return l2(out.as_array(), g["psi_out_f"] * (1 + k), float(g["dx"])) / (1 + k)

    with ThreadPoolExecutor(4) as pool:
        errs = list(pool.map(step, range(16)))
    assert max(errs) <= 1e-10, errs
