"""``FittingSolver`` with the reference's constructor and ``time_propagation`` entry point
(fitter.py:156-178, 491-571), running every sweep as batched GPU launches.

The reference drives one ``PropagationSolver`` per basis vector step by step from Python and
evaluates the field update in a callback (fitter.py:284-333, 650-811).  Here a whole sweep of all
basis vectors is one kernel launch (``qcontrol_b200.control``) with the field update inside the
kernel; this class only keeps the reference's outward behaviour: the same reporter objects are
created (``create_propagation_reporter("iter_{k}f/basis_{v}", nlevs)``), they receive
``print_time_point_prop`` for step 0 and for every step their ``mod_fileout`` / ``lmin`` select, and
``print_iter_point_fitter`` receives the same rows (iter, goal_close, E_tlist, t_list, Fsm, E_int, J, nt).

Supported task types: single_pot, filtering, trans_wo_control, intuitive_control,
local_control_population, local_control_projection, optimal_control_krotov,
optimal_control_unit_transform.  For local_control_projection the feedback of every step's overlaps
into the carrier frequency (fitter.py:598-623, 814-836) runs inside the sweep kernel.
"""
from __future__ import annotations

import math
import os
from types import SimpleNamespace

import numpy

from . import control, engine as _eng, functionals, math_base, phys_base
from .config import TaskRootConfiguration
from .phys_base import ExpectationValues, SigmaExpectationValues
from .propagation import PropagationSolver
from .psi_basis import Psi, PsiBasis

_TT = TaskRootConfiguration.TaskType
_NAMES = {_TT.SINGLE_POT: "single_pot", _TT.FILTERING: "filtering", _TT.TRANS_WO_CONTROL: "trans_wo_control",
          _TT.INTUITIVE_CONTROL: "intuitive_control", _TT.LOCAL_CONTROL_POPULATION: "local_control_population",
          _TT.LOCAL_CONTROL_PROJECTION: "local_control_projection", _TT.OPTIMAL_CONTROL_KROTOV: "optimal_control_krotov",
          _TT.OPTIMAL_CONTROL_UNIT_TRANSFORM: "optimal_control_unit_transform"}
_HAMIL = {_eng.HAMIL_NONTRIV: "ntriv", _eng.HAMIL_TWO_LEVELS: "two_levels", _eng.HAMIL_BH_X: "bh_x",
          _eng.HAMIL_BH_Z: "bh_z", _eng.HAMIL_QUBIT: "qubit"}


class FitterReporter:
    """Interface the control loop reports to (reporter.py:564-578 of the reference)."""

    def open(self):
        raise NotImplementedError()

    def close(self):
        raise NotImplementedError()

    def print_iter_point_fitter(self, iter, goal_close, E_tlist, t_list, Fsm, E_int, J, nt):
        raise NotImplementedError()

    def create_propagation_reporter(self, prop_id: str, nlvls: int):
        raise NotImplementedError()


def report_stride(rep) -> int:
    """The steps at which a propagation reporter must be called.  The reference calls every reporter at every step
    and lets each filter on ``l % mod == 0`` (reporter.py:110-164, tests/test_tools.py:189-194); the GPU path
    synchronises only on multiples of the returned stride.  For a fan-out reporter this is the GREATEST COMMON
    DIVISOR of its members' moduli -- every member still filters on its own modulus, so no record any of them would
    keep is skipped (the minimum would lose step 100 for members with moduli 100 and 30).  0 = no member ever
    keeps a step."""
    if hasattr(rep, "reps"):          # MultiplePropagationReporter
        g = 0
        for r in rep.reps:
            g = math.gcd(g, report_stride(r))
        return g
    for holder in (rep, getattr(rep, "conf", None), getattr(rep, "_conf", None)):
        if holder is None:
            continue
        for name in ("mod_fileout", "mod_plotout", "mod_update"):
            v = getattr(holder, name, None)
            if isinstance(v, int) and v > 0:
                return v
    return 1


class BatchObserver:
    """The reporter side of a batched sweep: for every run of a ``control`` batch one FitterReporter; at step 0
    and at every step the propagation reporters keep, the instrumentation of report_static / report_dynamic
    (propagation.py:184-321) is evaluated on the device for the whole batch and each run's reporters receive
    ``print_time_point_prop`` with the arguments the reference passes."""

    def __init__(self, batch, reporters):
        assert len(reporters) == batch.B
        self.batch, self.reporters = batch, list(reporters)
        self._open_id, self._reps = None, []

    def close(self):
        for per_run in self._reps:
            for r in per_run:
                r.close()
        self._reps = []

    def __call__(self, prop_id, l, direction, snaps):
        batch = self.batch
        pl = batch.plan
        p0 = batch.pbs[0]
        if prop_id != self._open_id:
            self.close()
            for rep, pb in zip(self.reporters, batch.pbs):
                per_run = []
                for v in range(pb.nb):
                    r = rep.create_propagation_reporter(os.path.join(prop_id, "basis_" + str(v)), pb.nlevs)
                    r.open()
                    per_run.append(r)
                self._reps.append(per_run)
            self._open_id = prop_id
            g = 0
            for per_run in self._reps:
                for r in per_run:
                    g = math.gcd(g, report_stride(r))
            batch.stride = g if g > 0 else 1 << 30
        fields = pl.get_fields() if l > 0 else None
        E_all, fm_all, Efull = [], [], numpy.zeros(batch.B, numpy.complex128)
        for b, pb in enumerate(batch.pbs):
            E = fields[b, min(l, pb.nt)] if l > 0 else self._E0_of(b, prop_id, direction)
            freq_mult = numpy.float64(1.0)
            if pb.task_type == "local_control_projection" and l > 0:      # (E_l, freq_mult_l) share one slot
                freq_mult, E = numpy.float64(numpy.imag(E)), numpy.float64(numpy.real(E))
            t = control.step_times(pb, direction)[min(l, pb.nt)]
            if pb.hf_hide and l > 0:
                exp_L = numpy.complex128(numpy.sqrt(complex(pb.laser_field_hf(freq_mult, t, pb.pcos, pb.w_list))))
                Efull[b] = E * exp_L * exp_L
            else:
                Efull[b] = E if l > 0 else 0.0
            E_all.append(E)
            fm_all.append(freq_mult)
        tab_all = pl.instrument(Efull, 2.0 * float(p0.m), snaps[0], snaps[1])       # [batch][nb][nlevs][20]
        state_all = pl.get_state()
        norms_all = pl.get_norms() if l > 0 else tab_all[:, :, :, 0]
        for b, pb in enumerate(batch.pbs):
            if l > pb.nt:                       # a shorter run of a ragged batch has already finished
                continue
            tab, state, norms, E = tab_all[b], state_all[b], norms_all[b], E_all[b]
            t = control.step_times(pb, direction)[l]
            shown = abs(E) if pb.ntriv == 1 else numpy.real(E)
            for v, rep in enumerate(self._reps[b]):
                tv = tab[v]
                c = lambda a, b_: tv[:, a] + 1j * tv[:, b_]
                moms = ExpectationValues(c(8, 9), c(10, 11), c(12, 13), c(14, 15))
                if pb.ntriv != 1:
                    moms.p = numpy.zeros(pb.nlevs, numpy.complex128)
                    moms.p2 = numpy.zeros(pb.nlevs, numpy.complex128)
                if pb.ntriv != 1 and pb.nlevs == 2:
                    cross = tv[0, 18] + 1j * tv[0, 19]
                    smoms = SigmaExpectationValues(2.0 * cross.real, 2.0 * cross.imag, complex(tv[0, 0] - tv[1, 0]))
                else:
                    smoms = SigmaExpectationValues(numpy.float64(0.0), numpy.float64(0.0), numpy.float64(0.0))
                ener, o0, of = c(2, 3), c(4, 5), c(6, 7)
                rep.print_time_point_prop(l, Psi.from_array(state[v]), t, pb.x, pb.np_, pb.nt, moms, smoms, ener,
                                          norms[v].astype(numpy.complex128), o0, of, [o0.sum(), of.sum()], ener.sum(),
                                          [tv[n, 16] for n in range(pb.nlevs)], [tv[n, 17] for n in range(pb.nlevs)], shown,
                                          fm_all[b])

    def _E0_of(self, b, prop_id, direction):
        """The field start() evaluates at l = 0 (propagation.py:360-361)."""
        pb = self.batch.pbs[b]
        it = int(prop_id.split("_")[1][:-1])
        if direction == control.BACKWARD:
            if pb.task_type == "optimal_control_unit_transform" and it > 0:
                return self.batch.plan.get_fields()[b, pb.nt]
            if pb.task_type == "optimal_control_krotov":
                return numpy.float64(0.0)
            return pb.laser_field(pb.E0, pb.T, pb.t0, pb.sigma) * (
                1.0 if pb.ntriv == 1 else pb.laser_field_hf(numpy.float64(1.0), pb.T, pb.pcos, pb.w_list))
        if pb.task_type == "optimal_control_krotov" and it > 0:
            return self.batch.plan.get_fields()[b, 0]
        env = pb.laser_field(pb.E0, numpy.float64(0.0), pb.t0, pb.sigma)
        return env if pb.hf_hide else env * pb.laser_field_hf(numpy.float64(1.0), numpy.float64(0.0), pb.pcos, pb.w_list)


class FittingSolver:
    class FitterDynamicState:
        def __init__(self, basis_length, levels_number):
            self.iter_step = 0
            self.Fsm = numpy.complex128(0)
            self.E_int = numpy.float64(0.0)
            self.J = numpy.float64(0.0)
            self.h_lambda = numpy.float64(0.0)
            self.goal_close_abs = numpy.float64(0.0)
            self.goal_close_vec = numpy.zeros(basis_length, numpy.complex128)
            self.E_tlist = []
            self.res = PropagationSolver.StepReaction.OK

    def __init__(self, conf_fitter, task_type, T, np, L, init_dir, ntriv, psi_init_basis: PsiBasis,
                 psi_goal_basis: PsiBasis, v, akx2, Fgoal, laser_field, laser_field_hf, F_type, aF_type, hamil2D,
                 reporter, _warning_collocation_points, _warning_time_steps):
        self._warning_collocation_points = _warning_collocation_points
        self._warning_time_steps = _warning_time_steps
        self.laser_field, self.laser_field_hf = laser_field, laser_field_hf
        self.F_type, self.aF_type, self.hamil2D = F_type, aF_type, hamil2D
        self.reporter = reporter
        self.task_type, self.T, self.np, self.L = task_type, T, np, L
        self.conf_fitter, self.init_dir, self.ntriv, self.Fgoal = conf_fitter, init_dir, ntriv, Fgoal
        self.psi_init_basis, self.psi_goal_basis = psi_init_basis, psi_goal_basis
        self.v, self.akx2 = v, akx2
        self.basis_length = len(psi_init_basis)
        self.levels_number = len(psi_init_basis.psis[0].f)
        self.dyn = None

    # ------------------------------------------------------------------------------------------
    def _functional_kind(self):
        """"sm" / "re" / "ss": which functional the F_type callable is (task_manager.py:162-207)."""
        return functionals.kind_of(self.F_type)

    def _carrier(self):
        """("exp", nu) when the high-frequency callable is exp(i 2 pi nu freq_mult t) (task_manager.py:73-86),
        found by probing it; ("other", nu_L) otherwise.  Only the in-kernel multiplier feedback needs to know."""
        import cmath
        cf, cp = self.conf_fitter, self.conf_fitter.propagation
        nu = getattr(getattr(self.laser_field_hf, "__self__", None), "nu", cp.nu_L)
        fm, t = numpy.float64(1.25), numpy.float64(self.T) * 0.37
        try:
            got = complex(self.laser_field_hf(fm, t, cf.pcos, cf.w_list))
        except Exception:
            return "other", nu
        want = cmath.exp(1j * 2.0 * math.pi * nu * fm * t)
        return ("exp" if got == want else "other"), nu

    def _problem(self, dx, x, t_step, t_list):
        cf, cp = self.conf_fitter, self.conf_fitter.propagation
        if self.task_type not in _NAMES:
            raise NotImplementedError("task type %s" % self.task_type)
        h = self.hamil2D
        mode = cf.h_lambda_mode
        return SimpleNamespace(
            task_type=_NAMES[self.task_type], hamil=_HAMIL[h.kind], ntriv=self.ntriv, np_=self.np,
            nlevs=self.levels_number, nb=self.basis_length, L=self.L, T=self.T, dx=dx, x=x, v=h.potentials(),
            vmin=h.minima(), akx2=numpy.asarray(self.akx2), psi0=self.psi_init_basis.as_array(),
            psif=self.psi_goal_basis.as_array(), U=h._U, W=h._W, delta=h._delta, m=cp.m, nch=cp.nch, nt=cp.nt, E0=cp.E0,
            t0=cp.t0, sigma=cp.sigma, nu_L=cp.nu_L, pcos=cf.pcos, w_list=cf.w_list, hf_hide=bool(cf.hf_hide),
            h_lambda=cf.h_lambda, h_lambda_mode=getattr(mode, "name", str(mode)).lower(), epsilon=cf.epsilon,
            iter_max=cf.iter_max, iter_mid_1=cf.iter_mid_1, iter_mid_2=cf.iter_mid_2, q=cf.q,
            impulses_number=cf.impulses_number, delay=cf.delay, F_type=self._functional_kind(), F_goal=self.Fgoal,
            k_E=cf.k_E, lamb=cf.lamb, pow_=cf.pow, nu=self._carrier()[1], carrier=self._carrier()[0],
            t_step=t_step, t_list=t_list, laser_field=self.laser_field, laser_field_hf=self.laser_field_hf)

    # ------------------------------------------------------------------------------------------
    def _close_reporters(self):
        if getattr(self, "_observer", None) is not None:
            self._observer.close()

    # ------------------------------------------------------------------------------------------
    def time_propagation(self, dx, x, t_step, t_list):
        """fitter.py:491-571"""
        pb = self._pb = self._problem(dx, x, t_step, t_list)
        self.dyn = FittingSolver.FitterDynamicState(self.basis_length, self.levels_number)
        grid = pb.hamil == "ntriv"
        batch = self._batch = (control.GridBatch if grid else control.NLevelBatch)([pb], ctx=phys_base.default_context())
        akx = None
        if grid:
            akx = numpy.ascontiguousarray((math_base.initak(pb.np_, dx, 1, pb.ntriv) *
                                           (phys_base.hart_to_cm / (-1j) / phys_base.dalt_to_au)).real)
        batch.plan.set_instr_grid(numpy.asarray(x, numpy.float64), akx)
        self._observer = batch.observer = BatchObserver(batch, [self.reporter])
        try:
            if pb.task_type == "optimal_control_krotov":
                hist = batch.run_krotov()
                if batch.errors[0]:                      # divergence guard (fitter.py:550-560)
                    self._emit(hist[0], t_list, pb)
                    raise ValueError(batch.errors[0])
            elif pb.task_type == "optimal_control_unit_transform":
                hist = batch.run_unit_transform()
                if batch.errors[0]:
                    self._emit(hist[0], t_list, pb)
                    raise ValueError(batch.errors[0])
            elif pb.task_type == "local_control_population":
                hist = batch.run_local_population()
            elif pb.task_type == "local_control_projection":
                hist = batch.run_local_projection()
                if batch.errors[0]:
                    raise RuntimeError(batch.errors[0])
            else:
                hist = batch.run_prescribed()
            self._emit(hist[0], t_list, pb)
        finally:
            self._close_reporters()
            batch.close()

    def _emit(self, rows, t_list, pb):
        for r in rows:
            self.reporter.print_iter_point_fitter(r.it, r.goal_close, list(r.E_tlist), t_list, numpy.complex128(r.Fsm),
                                                  numpy.float64(r.E_int), r.J, pb.nt)
        if rows:
            last = rows[-1]
            d = self.dyn
            d.iter_step, d.Fsm, d.E_int, d.J = last.it, numpy.complex128(last.Fsm), numpy.float64(last.E_int), last.J
            d.goal_close_abs = numpy.float64(last.goal_close)
            d.E_tlist = list(last.E_tlist)
