"""``PropagationSolver`` with the reference's constructor, ``start`` / ``step`` protocol, public
``stat`` / ``dyn`` / ``instr`` attributes, callbacks and reporter calls (propagation.py:17-484) --
with every array operation of a step on the GPU.

One ``step`` is ONE launch of the fused sweep kernel (rotating-frame transform, Newton polynomial,
renormalisation, back transform) for a batch of one, followed by the device-side instrumentation
(energy, overlaps, moments).  This class exists for call-by-call compatibility (it synchronises
with the host every step, exactly like the callbacks of the reference require); throughput comes
from the batched sweeps in ``qcontrol_b200.control`` / ``qcontrol_b200.fitter``.
"""
from __future__ import annotations

import cmath
import copy
import datetime
import math
from enum import Enum
from typing import Callable, Optional

import numpy

from . import engine as _eng
from . import math_base, phys_base
from .phys_base import ExpectationValues, SigmaExpectationValues, Session
from .psi_basis import Psi


class PropagationReporter:
    """Interface the stepper reports to (reporter.py:57-70 of the reference)."""

    def __init__(self, out_path: str = "", nlvls: int = 2):
        self._out_path = out_path
        self._nlvls = nlvls

    def open(self):
        raise NotImplementedError()

    def close(self):
        raise NotImplementedError()

    def print_time_point_prop(self, l, psi, t, x, np, nt, moms, smoms, ener, norm, overlp0, overlpf, overlp_tot,
                              ener_tot, psi_max_abs, psi_max_real, E, freq_mult):
        raise NotImplementedError()


class PropagationSolver:
    class StaticState:
        def __init__(self, psi0, psif, moms0, smoms0, cnorm0, cnormf, cener0, cenerf, overlp00, overlpf0,
                     dt=0.0, dx=0.0, x=numpy.empty(shape=0, dtype=numpy.float64), v=None, akx2=None):
            assert (psi0 is None and psif is None) or (psi0.f[0] is not psif.f[0] and psi0.f[1] is not psif.f[1]), \
                "A single array is passed twice (as psi0 and psif). Clone it!"
            self.psi0, self.psif = psi0, psif
            self.moms0, self.smoms0 = moms0, smoms0
            self.cnorm0, self.cnormf, self.cener0, self.cenerf = cnorm0, cnormf, cener0, cenerf
            self.overlp00, self.overlpf0 = overlp00, overlpf0
            self.dt, self.dx, self.x, self.v, self.akx2 = dt, dx, x, v, akx2

    class DynamicState:
        def __init__(self, l=0, t=0.0, psi: Psi = Psi(None), psi_omega: Psi = Psi(None), E=0.0,
                     freq_mult: numpy.float64 = 1.0, dir=None):
            assert (psi is None and psi_omega is None) or \
                   (psi.f[0] is not psi_omega.f[0] and psi.f[1] is not psi_omega.f[1]), \
                "A single array is passed twice (as psi and psi_omega). Clone it!"
            self.l, self.t, self.psi, self.psi_omega = l, t, psi, psi_omega
            self.E = E
            self.freq_mult = numpy.float64(freq_mult)
            self.dir = dir

    class InstrumentationOutputData:
        def __init__(self, moms, smoms, cnorm, psigc_psie, psigc_dv_psie, cener, E_full, overlp0, overlpf, emax, emin,
                     t_sc, time_before, time_after):
            self.moms, self.smoms, self.cnorm = moms, smoms, cnorm
            self.psigc_psie, self.psigc_dv_psie = psigc_psie, psigc_dv_psie
            self.cener, self.E_full, self.overlp0, self.overlpf = cener, E_full, overlp0, overlpf
            self.emax, self.emin, self.t_sc = emax, emin, t_sc
            self.time_before, self.time_after = time_before, time_after

    class StepReaction(Enum):
        OK = 0
        CORRECT = 1
        ITERATE = 2

    class Direction(Enum):
        FORWARD = 1
        BACKWARD = -1

    stat: Optional[StaticState]
    dyn: Optional[DynamicState]
    instr: Optional[InstrumentationOutputData]

    def __init__(self, T, np, L, _warning_collocation_points, _warning_time_steps, reporter, hamil2D,
                 laser_field_envelope, laser_field_hf, freq_multiplier: Callable, dynamic_state_factory, pcos, w_list,
                 mod_log, ntriv, hf_hide, conf_prop):
        self.milliseconds_full = 0.0
        self.T, self.np, self.L = T, np, L
        self._warning_collocation_points = _warning_collocation_points
        self._warning_time_steps = _warning_time_steps
        self.reporter = reporter
        self.hamil2D = hamil2D
        self.laser_field_envelope = laser_field_envelope
        self.laser_field_hf = laser_field_hf
        self.freq_multiplier = freq_multiplier
        self.dynamic_state_factory = dynamic_state_factory
        self.mod_log, self.ntriv, self.hf_hide = mod_log, ntriv, hf_hide
        self.pcos, self.w_list = pcos, w_list
        self.m, self.nch, self.nt = conf_prop.m, conf_prop.nch, conf_prop.nt
        self.E0, self.nu_L = conf_prop.E0, conf_prop.nu_L
        self.stat = self.dyn = self.instr = None
        self._s: Optional[Session] = None

    # ------------------------------------------------------------------------------------------
    def _instrument(self, psi: Psi, E_full):
        """cnorm, cener, overlaps, moments of ``psi`` -- one device call (qc_instrument)."""
        self._s.upload(psi)
        return self._table(self._s.plan.instrument(E_full, 2.0 * float(self.m), 0, 1)[0, 0])

    def _extremes(self):
        """propagation.py:371-375"""
        extr = []
        kmax = abs(self.stat.akx2[int(self.np / 2 - 1)])
        for n in range(len(self.stat.psi0.f)):
            extr.append(self.stat.v[n][1][0] + kmax)
            extr.append(self.stat.v[n][0])
        return extr

    def _resolution_check(self, emax, emin):
        nt_min = int(math.ceil((emax - emin) * self.T * phys_base.cm_to_erg / 2.0 / phys_base.Red_Planck_h))
        np_min = int(math.ceil(self.L * math.sqrt(2.0 * self.m * (emax - emin) * phys_base.dalt_to_au /
                                                  phys_base.hart_to_cm) / math.pi))
        if self.np < np_min and self._warning_collocation_points:
            self._warning_collocation_points(self.np, np_min)
        if self.nt < nt_min and self._warning_time_steps:
            self._warning_time_steps(self.nt, nt_min)

    def _shown_field(self):
        return abs(self.dyn.E) if self.ntriv == 1 else numpy.real(self.dyn.E)

    def report_static(self):
        """propagation.py:184-258"""
        extr = self._extremes()
        self._resolution_check(max(extr), min(extr))
        i0 = self._i0
        nl = len(self.stat.psi0.f)
        ener_tot = sum(self.stat.cener0[n] for n in range(nl))
        o0 = sum(self.stat.overlp00[n] for n in range(nl))
        of = sum(self.stat.overlpf0[n] for n in range(nl))
        self.reporter.print_time_point_prop(self.dyn.l, self.stat.psi0, self.dyn.t, self.stat.x, self.np, self.nt,
                                            self.stat.moms0, self.stat.smoms0, self.stat.cener0, self.stat.cnorm0,
                                            self.stat.overlp00, self.stat.overlpf0, [o0, of], ener_tot,
                                            i0["max_abs"], i0["max_real"], self._shown_field(), 1.0)

    def report_dynamic(self):
        """propagation.py:260-321"""
        ins, i = self.instr, self._i
        nl = len(self.stat.psi0.f)
        ener_tot = sum(ins.cener[n] for n in range(nl))
        o0 = sum(ins.overlp0[n] for n in range(nl))
        of = sum(ins.overlpf[n] for n in range(nl))
        span = ins.time_after - ins.time_before
        self.milliseconds_full += span.microseconds / 1000
        self.reporter.print_time_point_prop(self.dyn.l, self.dyn.psi, self.dyn.t, self.stat.x, self.np, self.nt, ins.moms,
                                            ins.smoms, ins.cener, ins.cnorm, ins.overlp0, ins.overlpf, [o0, of],
                                            ener_tot, i["max_abs"], i["max_real"], self._shown_field(),
                                            self.dyn.freq_mult)
        if self.dyn.l % self.mod_log == 0:
            self._resolution_check(ins.emax, ins.emin)

    # ------------------------------------------------------------------------------------------
    def start(self, v, akx2, dx, x, t_step, psi0: Psi, psif: Psi, dir: "PropagationSolver.Direction"):
        """propagation.py:323-364"""
        nl = len(psi0.f)
        self._s = Session(self.hamil2D, nl, self.nch, hf_hide=bool(self.hf_hide), dx=dx, x=x)
        pl = self._s.plan
        pl.snapshot_set(0, psi0.as_array()[None, None])
        pl.snapshot_set(1, psif.as_array()[None, None])
        self.stat = PropagationSolver.StaticState(psi0, psif, None, None, None, None, None, None, None, None,
                                                  dir.value * t_step, dx, x, v, akx2)
        inf = self._instrument(psif, 0.0)
        i0 = self._instrument(psi0, 0.0)
        self._i0 = i0
        st = self.stat
        st.moms0, st.smoms0 = i0["moms"], i0["smoms"]
        st.cnorm0, st.cener0, st.overlp00, st.overlpf0 = i0["cnorm"], i0["cener"], i0["overlp0"], i0["overlpf"]
        st.cnormf, st.cenerf = inf["cnorm"], inf["cener"]
        psi = copy.deepcopy(psi0)
        t0 = numpy.float64(0.0) if dir == PropagationSolver.Direction.FORWARD else self.T
        self.dyn = self.dynamic_state_factory(0, t0, psi, psi, numpy.float64(0.0), numpy.float64(1.0), dir)
        self.dyn.E = self.laser_field_envelope(self, self.stat, self.dyn)
        self.report_static()
        self.dyn.l = 1

    def step(self, t_start):
        """propagation.py:366-484: one time step, returns False after step nt."""
        time_before = datetime.datetime.now()
        s, pl, dyn, stat = self._s, self._s.plan, self.dyn, self.stat
        extr = self._extremes()
        dyn.t = stat.dt * dyn.l + t_start
        eL = self.nu_L * dyn.freq_mult * phys_base.Hz_to_cm / 2.0
        exp_L = numpy.complex128(1.0)
        if self.hf_hide:
            dyn.freq_mult = self.freq_multiplier(dyn, stat)
            exp_L = numpy.complex128(cmath.sqrt(self.laser_field_hf(dyn.freq_mult, dyn.t, self.pcos, self.w_list)))
            # callbacks (Krotov field updates) read dyn.psi_omega before the propagation
            dyn.psi_omega.f[0][:] = dyn.psi.f[0] / exp_L
            dyn.psi_omega.f[1][:] = dyn.psi.f[1] * exp_L
            ext = [extr[0] + self.E0 + eL, extr[1] - self.E0 + eL, extr[2] + self.E0 - eL, extr[3] - self.E0 - eL]
            emax, emin = max(ext), min(ext)
        else:
            emax, emin = max(extr), min(extr)
        t_sc = stat.dt * (emax - emin) * phys_base.cm_to_erg / 4.0 / phys_base.Red_Planck_h
        dyn.E = self.laser_field_envelope(self, stat, dyn)
        E_full = dyn.E * exp_L * exp_L if self.hf_hide else dyn.E
        # ---- the step itself: one fused kernel launch ---------------------------------------------
        s.set_poly(t_sc, emin, emax, eL)
        s.upload(dyn.psi)
        scal = numpy.zeros(2, numpy.complex128)
        scal[1] = dyn.E
        pl.set_E_base(scal)
        if self.hf_hide:
            scal2 = numpy.ones(2, numpy.complex128)
            scal2[1] = exp_L
            pl.set_exp_L(scal2, 0)
        pl.sweep(0, _eng.FIELD_PRESCRIBED, 1, 1, True, -1, -1)
        new = pl.get_state()[0, 0]
        cnorm = list(pl.get_norms()[0, 0].astype(numpy.complex128))
        psigc_psie = numpy.float64(0.0)
        psigc_dv_psie = numpy.float64(0.0)
        if self.hf_hide:
            dyn.psi.f[0] = new[0].copy()
            dyn.psi.f[1] = new[1].copy()
            dyn.psi_omega.f[0][:] = new[0] / exp_L
            dyn.psi_omega.f[1][:] = new[1] * exp_L
            psigc_psie = math_base.cprod(dyn.psi_omega.f[1], dyn.psi_omega.f[0], stat.dx, self.np)
            psigc_dv_psie = math_base.cprod3(dyn.psi_omega.f[1], stat.v[0][1] - stat.v[1][1], dyn.psi_omega.f[0],
                                             stat.dx, self.np)
        else:
            for n in range(len(new)):
                dyn.psi.f[n][...] = new[n]
        i = pl.instrument(E_full, 2.0 * float(self.m), 0, 1)
        self._i = self._table(i[0, 0])
        time_after = datetime.datetime.now()
        ii = self._i
        self.instr = PropagationSolver.InstrumentationOutputData(ii["moms"], ii["smoms"], cnorm, psigc_psie, psigc_dv_psie,
                                                                 ii["cener"], E_full, ii["overlp0"], ii["overlpf"], emax,
                                                                 emin, t_sc, time_before, time_after)
        self.report_dynamic()
        dyn.l += 1
        return dyn.l <= self.nt

    def _table(self, tab):
        c = lambda a, b: tab[:, a] + 1j * tab[:, b]
        nl = tab.shape[0]
        moms = ExpectationValues(c(8, 9), c(10, 11), c(12, 13), c(14, 15))
        if self.ntriv != 1:
            moms.p = numpy.zeros(nl, numpy.complex128)
            moms.p2 = numpy.zeros(nl, numpy.complex128)
        if self.ntriv != 1 and nl == 2:
            cross = tab[0, 18] + 1j * tab[0, 19]
            smoms = SigmaExpectationValues(2.0 * cross.real, 2.0 * cross.imag, complex(tab[0, 0] - tab[1, 0]))
        else:
            smoms = SigmaExpectationValues(numpy.float64(0.0), numpy.float64(0.0), numpy.float64(0.0))
        return dict(cnorm=c(0, 1), cener=c(2, 3), overlp0=c(4, 5), overlpf=c(6, 7), moms=moms, smoms=smoms,
                    max_abs=[tab[n, 16] for n in range(nl)], max_real=[tab[n, 17] for n in range(nl)])
