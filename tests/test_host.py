"""Host-side logic of the product package against the oracle / reference goldens (CPU only):
exact precompute (math_base), configuration DSL, task set-up, step scalars, batch enumeration."""
import ast
import contextlib
import io

import numpy as np
import pytest

from oracle import qc_oracle as qo

CONFS = ["fit_single_harm", "fit_single_morse", "fit_trans_woc", "fit_krotov_256", "fit_krotov_1024", "fit_ut_jx",
         "fit_ut_jx_4lvls", "fit_ut_jz", "fit_ut_s2s", "fit_pi_pulse", "fit_ut_jx_re", "fit_ut_jx_ss",
         "fit_ut_jx_maxwell_cos", "fit_ut_jx_const_sin"]


def _load(user):
    from qcontrol_b200 import config, task_manager
    with contextlib.redirect_stdout(io.StringIO()):
        c = config.TaskRootConfiguration()
        c.load(user)
        return c, task_manager.create(c)


def test_math_base_bit_exact(golden):
    """The uploaded interpolation tables are the reference's, bit for bit (math_base.py:153-188)."""
    from qcontrol_b200 import math_base as mb
    g = golden("kat_math_base")
    assert mb.reorder(8) == [0, 4, 2, 6, 1, 5, 3, 7] and mb.reorder(2) == [0, 1]
    for nch in (2, 4, 8, 64):
        assert mb.reorder(nch) == list(g["reorder_%d" % nch])
    for nch, t in ((4, 1.0), (8, 5.65e-3), (8, -5.65e-3), (64, 0.2040), (64, 3.737), (64, -3.737), (64, 20.55)):
        tag = "points_%d_%s" % (nch, repr(t).replace("-", "m").replace(".", "p"))
        xp, dv = mb.points(nch, t, qo.func)
        assert np.array_equal(xp, g[tag + "_xp"]) and np.array_equal(dv, g[tag + "_dv"])
        xp2, dv2 = mb.newton_table(nch, t)
        assert np.array_equal(xp2, xp) and np.array_equal(dv2, dv)
        xb, db = mb.newton_tables_batched(nch, [t, 0.5 * t])
        assert np.array_equal(xb, xp) and np.allclose(db[0], dv, rtol=0, atol=1e-18 + 1e-15 * np.abs(dv).max())
    assert np.array_equal(mb.initak(4, 0.2, 2, 1), g["initak_4"])
    assert np.array_equal(mb.initak(16, 0.3, 1, 1), g["initak_16_o1"])
    assert np.array_equal(mb.initak(8, 0.3, 2, -2), g["initak_8_triv"])
    s = np.array([1, 2, 3, 4])
    assert mb.cprod(s, s, 0.2, 4) == 6 and mb.cprod2(np.ones(4), s, 0.2, 4) == 2     # reference math_base_tests.py:6-16
    with pytest.raises(AssertionError):
        mb.cprod(s, s[:3], 0.2, 4)


@pytest.mark.parametrize("name", CONFS)
def test_task_setup_bit_exact(golden, name):
    """config.load + task_manager.create reproduce the reference's psi0 / psif / v / akx2 / grids exactly."""
    g = golden(name)
    c, tm = _load(ast.literal_eval(str(g["conf_json"])))
    assert np.array_equal(tm.psi0.as_array(), g["psi0"])
    if name != "fit_pi_pulse":                      # that test replaces the goal by hand (functional_tests.py:45-50)
        assert np.array_equal(tm.psif.as_array(), g["psif"])
    assert np.array_equal(np.stack([v[1] for v in tm.v]), g["v"])
    assert np.array_equal(np.array([float(v[0]) for v in tm.v]), g["vmin"])
    assert np.array_equal(tm.akx2, g["akx2"]) and np.array_equal(tm.x, g["x"])
    assert float(tm.dx) == float(g["dx"]) and float(tm.t_step) == float(g["t_step"])


def test_task_functions_kat(golden):
    """The product's envelopes, carriers (qcontrol_b200.task_manager), functionals and 'a' coefficients
    (qcontrol_b200.functionals, the one definition task_manager and control share) against the reference's own
    functions on the inputs of gen_golden.kat_task_inputs: bit-exact.  Covers the variants no shipped configuration
    uses: maxwell / const envelopes, cos / sin carriers, F_re / F_ss, a_re / a_ss (task_manager.py:24-268)."""
    from qcontrol_b200 import functionals, task_manager as tmod
    from qcontrol_b200.psi_basis import PsiBasis
    from tests.golden.gen_golden import kat_task_inputs
    g = golden("kat_task_functions")
    env_in, hf_in, gcs, bases = kat_task_inputs()
    L, H = tmod._LaserFields, tmod._LaserFieldsHighFrequencyPart
    for name, fn in (("zero", L.zero), ("const", L.const), ("gauss", L.laser_field_gauss),
                     ("sqrsin", L.laser_field_sqrsin), ("maxwell", L.laser_field_maxwell)):
        assert np.array_equal(np.array([fn(*a) for a in env_in], np.float64), g["env_" + name]), name
    for name, fn in (("exp", H.cexp), ("cos", H.cos), ("sin", H.sin), ("cos_set", H.cos_set), ("sin_set", H.sin_set)):
        assert np.array_equal(np.array([fn(*a) for a in hf_in], np.complex128), g["hf_" + name]), name
    for kind in functionals.KINDS:
        F_ref_style = getattr(tmod._F_type, "F_" + kind)
        a_ref_style = getattr(tmod._aF_type, "a_" + kind)
        assert functionals.kind_of(F_ref_style) == kind
        for k, gc in enumerate(gcs):
            want = g["F_%s_%d" % (kind, k)]
            assert np.complex128(F_ref_style(gc, len(gc))) == want and functionals.F_value(kind, gc, len(gc)) == want
            assert abs(functionals.F_batch(kind, gc[None, :])[0] - want.real) <= 1e-14 * max(1.0, abs(want))
        for k, (psi, chi, dx) in enumerate(bases):
            want = g["a_%s_%d" % (kind, k)]
            nb, nl, n = psi.shape
            assert np.array_equal(a_ref_style(PsiBasis.from_array(psi), PsiBasis.from_array(chi), dx, nb, nl, n), want)
            assert np.array_equal(functionals.a_value(kind, psi, chi, dx), want)
            assert np.allclose(functionals.a_batch(kind, psi[None], chi[None], np.array([dx]))[0], want, rtol=1e-13, atol=1e-13)
    # the vectorised set-up path of big batches (control._env_vec / _hf_vec) is the same formula up to libm rounding
    from qcontrol_b200 import control
    for name in ("zero", "const", "gauss", "sqrsin", "maxwell"):
        got = np.array([control._env_vec(name, a[0], np.array([a[1]]), a[2], a[3])[0] for a in env_in])
        assert np.allclose(got, g["env_" + name], rtol=1e-13, atol=0), name
    for name in ("exp", "cos", "sin", "cos_set", "sin_set"):
        got = np.array([control._hf_vec(name, a[0] * a[1], np.array([a[2]]), a[3], a[4])[0] for a in hf_in])
        assert np.allclose(got, g["hf_" + name], rtol=1e-11, atol=1e-12), name


def test_config_semantics():
    from qcontrol_b200.config import TaskRootConfiguration as T
    c = T()
    with contextlib.redirect_stdout(io.StringIO()) as out:
        c.load({"task_type": "optimal_control_krotov", "np": 256, "T": 1e-15, "unknown_key": 3,
                "fitter": {"hf_hide": False, "propagation": {"nt": 10}}})
    assert c.task_type == T.TaskType.OPTIMAL_CONTROL_KROTOV and c.np == 256 and isinstance(c.T, np.float64)
    assert c.fitter.propagation.nt == 10 and c.fitter.propagation.nch == 64 and c.fitter.hf_hide == 0
    assert "Parameter 'L' hasn't been provided" in out.getvalue()
    assert T.TaskType.from_int(7) == T.TaskType.OPTIMAL_CONTROL_UNIT_TRANSFORM
    with pytest.raises(KeyError):
        T.TaskType.from_str("nonsense")
    with pytest.raises(AttributeError):
        c.no_such_field
    c.fitter.w_list = [1.0]
    assert c.fitter.w_list == [1.0]


def test_factory_validation_errors():
    base = {"task_type": "single_pot", "pot_type": "morse", "wf_type": "morse", "hamil_type": "ntriv", "np": 256,
            "fitter": {"propagation": {"nt": 4}}}
    for patch, exc in (({"nlevs": 3}, RuntimeError), ({"nb": 2}, RuntimeError), ({"hamil_type": "two_levels", "nb": 2,
                                                                                   "fitter": {"hf_hide": False}}, RuntimeError),
                       ({"init_guess_hf": "cos"}, RuntimeError)):
        user = dict(base)
        user.update(patch)
        with pytest.raises(exc):
            _load(user)
    with pytest.raises(RuntimeError):
        _load({"task_type": "optimal_control_unit_transform", "pot_type": "none", "np": 1, "hamil_type": "BH_model",
               "fitter": {"hf_hide": True}})


@pytest.mark.parametrize("name", ["harm", "morse", "krotov"])
def test_energy_window_matches_reference(golden, name):
    """emax / emin / t_sc of PropagationSolver.step (propagation.py:371-407) as recorded from the reference."""
    from qcontrol_b200 import control
    from tests.golden.gen_golden import conf_krotov, conf_single_harm, conf_single_morse
    g = golden("one_step_" + name)
    user = {"harm": conf_single_harm(2000), "morse": conf_single_morse(2000), "krotov": conf_krotov(2000, 1)}[name]
    pb = qo.make_problem(user)
    emax, emin, t_sc, eL = control.energy_window(pb)
    assert (float(emax), float(emin), float(t_sc), float(eL)) == (float(g["emax"]), float(g["emin"]), float(g["t_sc"]),
                                                                  float(g["eL"]))
    _, _, t_b, _ = control.energy_window(pb, control.BACKWARD)
    assert float(t_b) == -float(t_sc)


def test_report_stride_is_the_gcd_of_the_member_moduli():
    """A fan-out reporter whose members keep every 100th and every 30th step must be visited at every multiple of
    gcd = 10 -- with the minimum (30) step 100 would never be seen."""
    from types import SimpleNamespace as NS
    from qcontrol_b200 import fitter
    member = lambda mod: NS(conf=NS(mod_fileout=mod))
    assert fitter.report_stride(member(40)) == 40
    assert fitter.report_stride(NS(reps=[member(100), member(30)])) == 10
    assert fitter.report_stride(NS(reps=[member(100), NS(reps=[member(75)])])) == 25
    assert fitter.report_stride(NS(reps=[])) == 0            # nobody keeps anything


def test_memory_sized_sub_batches_host_logic():
    """newcheb.memory_chunks: runs longest first, a sub-batch per memory budget, every run exactly once; the
    footprint bound of the shipped Krotov input (nt = 230000, np = 1024) is 2 x 3.77 GB of trajectories + state."""
    from qcontrol_b200 import engine, newcheb
    from tests.golden.gen_golden import conf_krotov, conf_single_morse
    confs = []
    for nt in (100, 400, 250, 400, 50, 300):
        c, _ = _load(conf_krotov(nt, 2, np_=1024))
        confs.append(c)
    one = engine.plan_footprint(engine.HAMIL_NONTRIV, 1024, 2, 1, 64, 1, 400, True, True)
    two = engine.plan_footprint(engine.HAMIL_NONTRIV, 1024, 2, 1, 64, 2, 400, True, True)
    assert two > one > 2 * 401 * 1024 * 16
    chunks = newcheb.memory_chunks(confs, list(range(6)), free_bytes=two / 0.8 + 1, fraction=0.8)
    assert sorted(i for c in chunks for i in c) == list(range(6))
    assert chunks[0] == [1, 3] and all(len(c) <= 3 for c in chunks) and len(chunks) >= 3
    assert newcheb.memory_chunks(confs, list(range(6)), free_bytes=1e12) == [[1, 3, 5, 2, 0, 4]]      # longest first
    with pytest.raises(MemoryError):
        newcheb.memory_chunks(confs, [1], free_bytes=one / 2)
    shipped = engine.plan_footprint(engine.HAMIL_NONTRIV, 1024, 2, 1, 64, 1, 230000, True, True)
    assert 7.5e9 < shipped < 7.8e9
    # a prescribed-field task has no trajectory buffers: the shipped filtering length costs megabytes, not 52 GB
    c, _ = _load(conf_single_morse(40))
    c.fitter.propagation.nt = 800000
    assert newcheb.memory_chunks([c], [0], free_bytes=1e9) == [[0]]


def test_substitution_expansion_order():
    """@SUBST:LIST expansion: Cartesian product, first list fastest (json_substitutions.py:70-99);
    the shipped 25 x 16 T-sweep expands to 400 variants."""
    from qcontrol_b200 import batch
    t = {"run_id": {"@SUBST:LIST": ["a", "b", "c"]}, "T": {"@SUBST:LIST": [1.0, 2.0]}, "fitter": {"x": [1, {"@SUBST:LIST": [7, 8]}]}}
    v = batch.expand_substitutions(t)
    assert len(v) == 12
    assert [(d["run_id"], d["T"], d["fitter"]["x"][1]) for d in v[:4]] == [("a", 1.0, 7), ("b", 1.0, 7), ("c", 1.0, 7), ("a", 2.0, 7)]
    assert v[-1] == {"run_id": "c", "T": 2.0, "fitter": {"x": [1, 8]}}
    assert batch.expand_substitutions({"a": 1}) == [{"a": 1}]
    assert batch.expand_substitutions({"a": {"@SUBST:LIST": []}}) == []
    sweep = {"run_id": {"@SUBST:LIST": ["ut/run%d" % i for i in range(1, 26)]}, "T": {"@SUBST:LIST": [2.5e-13] + [1e-12] * 15}}
    assert len(batch.expand_substitutions(sweep)) == 400


def test_auto_rules_and_seeded_w_list():
    from qcontrol_b200 import batch
    from tests.golden.gen_golden import conf_ut_jx
    user = conf_ut_jx(1, 1, T=11.8e-12)
    user["nt_auto"] = True
    user["fitter"]["w_list"] = []
    c, _ = _load(user)
    batch.apply_auto_rules(c, run_index=3)
    assert c.fitter.propagation.nt == 5900 and float(c.fitter.propagation.sigma) == 2.0 * 11.8e-12
    assert np.array_equal(c.fitter.w_list, np.random.default_rng(1003).uniform(-2.0, 2.0, 3))


def test_lpt_sharding_is_balanced_and_deterministic():
    from qcontrol_b200 import batch
    nts = [125, 5900, 6700, 7000, 7500, 8300, 8400, 8500, 8800, 8900, 9000, 9100, 9200, 9300, 9400, 9500] * 25
    shards = batch.lpt_shards(nts, 8)
    assert sorted(i for s in shards for i in s) == list(range(400))
    loads = [sum(nts[i] for i in s) for s in shards]
    assert max(loads) / min(loads) < 1.01
    assert shards == batch.lpt_shards(nts, 8)
    assert batch.lpt_shards([3.0], 4) == [[0], [], [], []]


@pytest.mark.parametrize("flag", ["tables", "tables_iter"])
def test_table_reporters_write_the_reference_formats(golden, tmp_path, flag):
    """tab_iter / tab_iter_E / tab_abs / tab_real / tab_tvals / tab_tvals_fit files, byte for byte
    what the reference's reporters (reporter.py:73-164, 581-632, 790-834) write for the same inputs."""
    import os
    from qcontrol_b200 import config, reporter
    from qcontrol_b200.phys_base import ExpectationValues, SigmaExpectationValues
    from qcontrol_b200.psi_basis import Psi
    import tests.golden.gen_golden as gg
    g = golden("report_tables")
    x, pts, iters = gg.report_inputs()
    out = str(tmp_path / "out")
    rep_json = {"fitter": {"out_path": out, "plotting_flag": flag, "table.imin": -1, "table.imod_fileout": 1,
                           "table.tab_iter": "tab_iter.csv", "table.tab_iter_E": "tab_iter_E.csv", "plot.imin": -1,
                           "propagation": {"table.lmin": 2, "table.mod_fileout": 2, "table.tab_abs": "tab_abs_{level}.csv",
                                           "table.tab_real": "tab_real_{level}.csv", "table.tab_tvals": "tab_tvals_{level}.csv",
                                           "table.tab_tvals_fit": "tab_tvals_fit.csv", "plot.lmin": 0}}}
    with contextlib.redirect_stdout(io.StringIO()):
        ct = config.ReportTableRootConfiguration()
        ct.load(rep_json)
        cp = config.ReportPlotRootConfiguration()
        cp.load(rep_json)
    assert ct.fitter.propagation.mod_fileout == 2 and ct.fitter.imin == -1 and cp.fitter.imin == -1
    fr = reporter.MultipleFitterReporter(conf_rep_table=ct.fitter, conf_rep_plot=cp.fitter)
    fr.open()
    pr = fr.create_propagation_reporter(os.path.join("iter_0f", "basis_0"), 2)
    pr.open()
    for p in pts:
        pr.print_time_point_prop(p["l"], Psi.from_array(p["psi"]), p["t"], x, 8, 4,
                                 ExpectationValues(p["momx"], p["momx2"], p["momp"], p["momp2"]),
                                 SigmaExpectationValues(p["smoms"][0], p["smoms"][1], p["smoms"][2]), p["ener"], p["norm"],
                                 p["overlp0"], p["overlpf"], p["overlp_tot"], p["ener_tot"], p["psi_max_abs"],
                                 p["psi_max_real"], p["E"], p["freq_mult"])
    pr.close()
    for it, gc, F, Ei, J, El, tl in iters:
        fr.print_iter_point_fitter(it, gc, list(El), tl, F, Ei, J, 4)
    fr.close()
    want = {k.split("::", 1)[1]: str(g[k]) for k in g if k.startswith(flag + "::")}
    got = {}
    for root, _, files in os.walk(out):
        for fn in files:
            with open(os.path.join(root, fn)) as f:
                got[os.path.relpath(os.path.join(root, fn), out).replace(os.sep, "/")] = f.read()
    assert sorted(got) == sorted(want)
    for k in want:
        assert got[k] == want[k], k


def test_shared_reciprocal_quotient_is_correctly_rounded():
    """The kernels divide many values by one divisor (renormalisation psi / sqrt(norm), the rotating frame, the
    field law / h_lambda) as rb = 1 / b once and  q = a rb,  r = fma(-q, b, a),  fma(r, rb, q)  per value
    (div_rn_by in csrc/qc_fft.cuh).  This restates the three steps with exact rational arithmetic for the two
    fused operations and checks that the result is the correctly rounded a / b the reference computes -- the
    reason the change of r01k left every result bit-identical."""
    from fractions import Fraction
    rng = np.random.default_rng(2024)
    n = 40000
    a = rng.standard_normal(n) * 10.0 ** rng.integers(-30, 3, n)
    b = np.concatenate([rng.uniform(0.5, 2.0, n // 2), 10.0 ** rng.uniform(-6, 3, n - n // 2)])
    a[:10] = 0.0
    bad = 0
    for x, y in zip(a.tolist(), b.tolist()):
        rb = 1.0 / y
        q = x * rb
        r = float(Fraction(x) - Fraction(q) * Fraction(y))          # fma(-q, y, x): one rounding
        q2 = float(Fraction(q) + Fraction(r) * Fraction(rb))        # fma(r, rb, q): one rounding
        bad += q2 != x / y
    assert bad == 0


def test_pipeline_pieces_cut_whole_waves():
    """newcheb.pipeline_pieces: a big grid sub-batch becomes pieces of two waves of CTAs (so that the set-up of the next
    piece can run beside the propagation of the current one), every run exactly once and in order; small batches and
    N-level batches stay whole."""
    from qcontrol_b200 import newcheb
    chunk = list(range(1024))
    pieces = newcheb.pipeline_pieces(chunk, 2048, 148, True)
    assert [len(p) for p in pieces] == [296, 296, 296, 136]
    assert sum(pieces, []) == chunk
    assert [len(p) for p in newcheb.pipeline_pieces(chunk, 1024, 148, True)] == [592, 432]
    assert newcheb.pipeline_pieces(chunk, 512, 148, True) == [chunk]             # one piece of 1184 holds them all
    assert [len(p) for p in newcheb.pipeline_pieces(chunk, 4096, 148, True)] == [148] * 6 + [136]
    assert newcheb.pipeline_pieces(chunk, 1, 148, False) == [chunk]              # kernel B: as many runs per launch as possible
    assert [len(p) for p in newcheb.pipeline_pieces(list(range(500)), 2048, 148, True)] == [296, 204]
    assert newcheb.pipeline_pieces(list(range(296)), 2048, 148, True) == [list(range(296))]
