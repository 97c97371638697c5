"""Batched control loops on the GPU engine.

``GridBatch`` runs many independent two-level FFT-grid problems (prescribed
field or Krotov optimal control) and ``NLevelBatch`` many N-level
unitary-transformation problems, each batch as ONE plan: all sweeps of all runs
are single kernel launches, the iteration logic stays on the host.  These
classes restate the sweep bookkeeping of ``FittingSolver`` (fitter.py:213-571)
for a whole batch; per run they produce the same iteration rows
(iter, goal_close, Fsm, E_int, J) and field lists E_tlist.

A "problem" is any object with the attributes produced by
``qcontrol_b200.task_manager.create`` (psi0, psif, v, vmin, akx2, dx, T, nt,
nch, E0, t0, sigma, nu_L, hf_hide, h_lambda, ... and the two callables
``laser_field`` / ``laser_field_hf``).
"""
from __future__ import annotations

import cmath
import math
import os
import time
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

from . import engine as eng
from . import math_base
from .functionals import F_batch, F_value as _F_value, a_batch, a_value as _a_value
from .phys_consts import CM_TO_ERG, HZ_TO_CM, RED_PLANCK_H

FORWARD, BACKWARD = 1, -1

_HAMIL_CODE = {"ntriv": eng.HAMIL_NONTRIV, "two_levels": eng.HAMIL_TWO_LEVELS, "bh_x": eng.HAMIL_BH_X,
               "bh_z": eng.HAMIL_BH_Z, "qubit": eng.HAMIL_QUBIT}


@dataclass
class IterRow:
    it: int
    goal_close: float
    Fsm: float
    E_int: float
    J: float
    E_tlist: np.ndarray


def energy_window(pb, direction=FORWARD, freq_mult=1.0):
    """emax, emin, t_sc, eL of PropagationSolver.step (propagation.py:371-407)."""
    extr = []
    kmax = abs(pb.akx2[int(pb.np_ / 2 - 1)])
    for n in range(pb.nlevs):
        extr.append(pb.v[n][0] + kmax)
        extr.append(pb.vmin[n])
    eL = pb.nu_L * freq_mult * HZ_TO_CM / 2.0
    if pb.hf_hide:
        ext = [extr[0] + pb.E0 + eL, extr[1] - pb.E0 + eL, extr[2] + pb.E0 - eL, extr[3] - pb.E0 - eL]
        emax, emin = max(ext), min(ext)
    else:
        emax, emin = max(extr), min(extr)
    dt = direction * pb.t_step
    t_sc = dt * (emax - emin) * CM_TO_ERG / 4.0 / RED_PLANCK_H
    return emax, emin, t_sc, eL


def step_times(pb, direction):
    """dyn.t for l = 0..nt (propagation.py:354-357, 377)."""
    nt = pb.nt
    if direction == FORWARD:
        t0 = np.float64(0.0)
        dt = pb.t_step
    else:
        t0 = pb.T
        dt = -pb.t_step
    return [t0] + [dt * l + t0 for l in range(1, nt + 1)]


# Vectorised envelope / carrier evaluation for LARGE batches.  The scalar callables of a problem
# (task_manager.py:18-159, Python ``math``) stay the source of truth and are used whenever the batch is
# small enough (every parity test); beyond VECTOR_SETUP_POINTS table entries the same formulas are
# evaluated with numpy over the whole time grid, which may differ from libm in the last ulp.
VECTOR_SETUP_POINTS = 200000


def _env_vec(name, E0, t, t0, sigma):
    if name == "zero":
        return np.zeros_like(t)
    if name == "const":
        return np.full_like(t, E0)
    if name == "gauss":
        return E0 * np.exp(-(t - t0) * (t - t0) / 2.0 / sigma / sigma)
    if name == "sqrsin":
        s = np.sin(2.0 * math.pi * (t - t0) / sigma)
        return E0 * s * s
    if name == "maxwell":
        return E0 * t * t * np.exp(-(t - t0) * (t - t0) / 2.0 / sigma / sigma) / sigma / sigma
    raise KeyError(name)


def _hf_vec(name, nu, t, pcos, w):
    base = 2.0 * math.pi * nu * 1.0 * t
    if name == "exp":
        return np.exp(1j * base)
    if name == "cos":
        return np.cos(base * pcos).astype(np.complex128)
    if name == "sin":
        return np.sin(base * pcos).astype(np.complex128)
    if name == "cos_set":
        acc = w[0] * np.cos(base)
        k = -1
        for q in range(2, math.floor(pcos) + 1):
            k += 2
            acc = acc + w[k] * np.cos(base * q) + w[k + 1] * np.cos(base / q)
        return acc.astype(np.complex128)
    if name == "sin_set":
        acc = w[0] * np.sin(base)
        for q in range(1, math.floor(pcos) + 1):
            acc = acc + w[q] * np.sin(base * (2 * q + 1))
        return acc.astype(np.complex128)
    raise KeyError(name)


def field_tables(p, direction, half_step=False, vectorised=False):
    """(envelope[l], carrier[l]) for l = 0..nt at t_l (or t_l - |dt|/2 for l >= 1 when half_step)."""
    ts = step_times(p, direction)
    if vectorised and getattr(p, "envelope", None) and getattr(p, "carrier", None):
        t = np.array(ts, np.float64)
        if half_step:
            t[1:] = t[1:] - abs(p.t_step) / 2.0
        return (np.asarray(_env_vec(p.envelope, p.E0, t, p.t0, p.sigma), np.float64),
                _hf_vec(p.carrier, p.nu, t, p.pcos, p.w_list))
    env, hf = [], []
    for k, t in enumerate(ts):
        th = t - (abs(p.t_step) / 2.0) if (half_step and k > 0) else t
        env.append(p.laser_field(p.E0, th, p.t0, p.sigma))
        hf.append(p.laser_field_hf(np.float64(1.0), th, p.pcos, p.w_list))
    return np.array(env, np.float64), np.array(hf, np.complex128)


def _shown_fields(Et):
    """|E| where E is complex, Re E otherwise: the list handed to print_iter_point_fitter (fitter.py:479-484)."""
    Et = np.asarray(Et, np.complex128)
    return np.where(Et.imag != 0.0, np.abs(Et), Et.real)


def field_integral(Et, t_step, exact=True):
    """E_int = sum_{l>=1} |E_l|^2 * dt[fs] (fitter.py:338-341).  exact: the reference's scalar accumulation order
    (bit-identical E_int); big batches pass exact=False and get the same sum from one numpy reduction."""
    Et = np.asarray(Et)
    if exact and Et.shape[0] <= 4096:
        acc = np.float64(0.0)
        for k in range(1, Et.shape[0]):
            acc += Et[k] * Et[k].conjugate() * (t_step * 1e15)
        return acc
    return np.complex128(np.sum(Et[1:] * np.conjugate(Et[1:])) * (t_step * 1e15))


TRAJ_TASKS = ("optimal_control_krotov", "optimal_control_unit_transform")


_EXPL_MEMO = {}      # (direction, carrier, nu, T, t_step, nt, pcos, w_list) -> exp_L[l] of GridBatch._upload_static


class _BatchBase:
    def __init__(self, problems: Sequence, ctx: eng.Context = None, device: int = 0, store_traj=None, poly_tol=None):
        """store_traj: allocate the two trajectory buffers [batch][nt_max+1][np] (grid) -- None = only for the task
        types whose field update reads the previous opposite sweep (Krotov, unitary transformation); pass True to
        record a prescribed-field sweep with ``propagate(record=True)``.  A filtering run of the shipped length
        (nt = 800000, np = 2048) would otherwise allocate 2 x 26 GB it never touches."""
        self.pbs = list(problems)
        p0 = self.pbs[0]
        if store_traj is None:
            store_traj = getattr(p0, "task_type", "") in TRAJ_TASKS
        # opt-in truncation of the interpolation series (qc_set_poly_orders): a per-step error bound, e.g. 1e-11
        if poly_tol is None and os.environ.get("QC_POLY_TOL"):
            poly_tol = float(os.environ["QC_POLY_TOL"])
        self.poly_tol = poly_tol
        self.B = len(self.pbs)
        self.ctx = ctx or eng.Context(device)
        for p in self.pbs:
            assert (p.np_, p.nlevs, p.nb, p.nch, p.hamil, p.hf_hide) == \
                   (p0.np_, p0.nlevs, p0.nb, p0.nch, p0.hamil, p0.hf_hide), "a batch shares one plan shape"
        self.nt = np.array([p.nt for p in self.pbs])
        self.nt_max = int(self.nt.max())
        _t = [time.perf_counter()]
        self.plan = eng.Plan(self.ctx, _HAMIL_CODE[p0.hamil], p0.np_, p0.nlevs, p0.nb, p0.nch, self.B, self.nt_max,
                             hf_hide=p0.hf_hide, store_traj=store_traj)
        _t.append(time.perf_counter())
        self.psi0 = np.stack([p.psi0 for p in self.pbs])
        self.psif = np.stack([p.psif for p in self.pbs])
        self.dx = np.array([float(p.dx) for p in self.pbs])
        self.plan.snapshot_set(0, self.psi0)      # SNAP_INIT
        self.plan.snapshot_set(1, self.psif)      # SNAP_GOAL
        _t.append(time.perf_counter())
        self._upload_static()
        _t.append(time.perf_counter())
        self._poly_dir = {}
        for slot, d in ((0, FORWARD), (1, BACKWARD)):
            self._upload_poly(slot, d)
        _t.append(time.perf_counter())
        if os.environ.get("QC_PROFILE_SETUP"):
            print("batch set-up: plan %.3f s, snapshots %.3f s, static + exp_L %.3f s, polynomials %.3f s (%d runs)"
                  % (_t[1] - _t[0], _t[2] - _t[1], _t[3] - _t[2], _t[4] - _t[3], self.B), flush=True)

    def _upload_poly(self, slot, direction):
        emins, emaxs, tscs, eLs = [], [], [], []
        for p in self.pbs:
            emax, emin, t_sc, eL = energy_window(p, direction)
            emins.append(emin); emaxs.append(emax); tscs.append(t_sc); eLs.append(eL)
        nch = self.pbs[0].nch
        if len(set(float(t) for t in tscs)) > 64:
            xp, dvs = math_base.newton_tables_batched(nch, tscs)
        else:
            tabs = [math_base.newton_table(nch, t) for t in tscs]
            xp, dvs = tabs[0][0], np.stack([t[1] for t in tabs])
        self.plan.set_poly(slot, xp, dvs, np.array(emins), np.array(emaxs), np.array(tscs), np.array(eLs))
        self._poly_dir[direction] = slot
        if self.poly_tol:
            self.poly_orders = getattr(self, "poly_orders", {})
            self.poly_orders[direction] = math_base.truncation_order(nch, dvs, self.poly_tol)
            self.plan.set_poly_orders(slot, self.poly_orders[direction])

    on_last_sweep = None  # callable() invoked once the last sweep of a control loop has been launched (run_krotov)
    time_sweeps = False  # accumulate CUDA-event time of the sweep launches in sweep_ms (synchronises)
    sweep_ms = 0.0
    observer = None      # callable(prop_id, l, direction, snapshots) invoked at l = 0 and every `stride` steps
    stride = 1

    def _sweep(self, prop_id, direction, snaps, slot, mode, l_first, nsteps, tr, tw, **kw):
        """One sweep launch, or -- when an observer is attached -- the same sweep cut into chunks that end
        on the steps the observer wants to see (the reporters of the reference are called every step and
        keep every mod-th one, propagation.py:291-299; here only those steps cost a host round trip)."""
        pl = self.plan
        if self.observer is None:
            if self.time_sweeps:
                self.ctx.timer_start()
            pl.sweep(slot, mode, l_first, nsteps, True, tr, tw, **kw)
            if self.time_sweeps:
                self.sweep_ms += self.ctx.timer_stop()
            return
        end = l_first + nsteps
        l = l_first
        if l == 0:
            pl.sweep(slot, mode, 0, 1, True, tr, tw, **kw)
            l = 1
        self.observer(prop_id, 0, direction, snaps)
        stride = max(1, int(self.stride))
        while l < end:
            nxt = min(end, ((l - 1) // stride + 1) * stride + 1)
            pl.sweep(slot, mode, l, nxt - l, True, tr, tw, **kw)
            l = nxt
            if (l - 1) % stride == 0:
                self.observer(prop_id, l - 1, direction, snaps)

    def _pad(self, rows, dtype=np.complex128):
        out = np.zeros((self.B, self.nt_max + 1), dtype)
        for b, r in enumerate(rows):
            out[b, :len(r)] = r
        return out

    def close(self):
        self.plan.close()


class GridBatch(_BatchBase):
    """Two-level FFT-grid runs: prescribed-field propagation and Krotov optimal control."""

    def _upload_static(self):
        p0 = self.pbs[0]
        self.plan.set_static(np.stack([p.v for p in self.pbs]), np.ascontiguousarray(p0.akx2.real), None, self.dx)
        for p in self.pbs:
            assert np.array_equal(p.akx2, p0.akx2), "one kinetic multiplier per batch"
        if p0.hf_hide:
            # exp_L[l] = csqrt(hf(t_l)) with the reference's scalar expressions (branch cut of cmath.sqrt).
            # The table depends on the time grid and the carrier only, which most runs of a batch share (the
            # variants of a template differ in a scalar or two): one evaluation per distinct (grid, carrier)
            # the sub-batches of a big template share them too: a small memo across batches (_EXPL_MEMO)
            for slot, d in ((0, FORWARD), (1, BACKWARD)):
                table = np.zeros((self.B, self.nt_max + 1), np.complex128)
                for b, p in enumerate(self.pbs):
                    key = None
                    if getattr(p, "carrier", None) and getattr(p, "nu", None) is not None:
                        # the exp carrier reads neither pcos nor w_list, cos / sin read pcos, the sets both
                        # (task_manager.py:68-159)
                        sets = p.carrier in ("cos_set", "sin_set")
                        key = (d, p.carrier, float(p.nu), float(p.T), float(p.t_step), int(p.nt),
                               0.0 if p.carrier == "exp" else float(p.pcos),
                               tuple(float(w) for w in p.w_list) if sets else ())
                    row = _EXPL_MEMO.get(key) if key is not None else None
                    if row is None:
                        row = np.array([np.complex128(cmath.sqrt(p.laser_field_hf(np.float64(1.0), t, p.pcos, p.w_list)))
                                        for t in step_times(p, d)], np.complex128)
                        if key is not None:
                            if len(_EXPL_MEMO) >= 64:
                                _EXPL_MEMO.clear()
                            _EXPL_MEMO[key] = row
                    table[b, :len(row)] = row
                self.plan.set_exp_L(table, slot)
        self.plan.set_control(np.array([float(p.h_lambda) for p in self.pbs]), self.nt.astype(np.float64))

    def _vectorised(self):
        """Big batches evaluate their per-step tables with numpy over the whole time grid instead of the reference's
        scalar expressions (last-ulp differences in libm calls; every parity test is below the threshold)."""
        return int(self.nt.sum()) > VECTOR_SETUP_POINTS

    def _prescribed_field(self, p, direction):
        """LaserFieldEnvelope / ...Backward for tasks without feedback (fitter.py:650-670, 809)."""
        key = (id(p), direction)
        cache = self.__dict__.setdefault("_presc_cache", {})
        if key not in cache:
            cache[key] = self._prescribed_field_eval(p, direction)
        return cache[key]

    def _prescribed_field_eval(self, p, direction):
        if self._vectorised() and getattr(p, "envelope", None) and getattr(p, "carrier", None):
            t = np.array(step_times(p, direction), np.float64)
            E = np.asarray(_env_vec(p.envelope, p.E0, t, p.t0, p.sigma), np.float64)
            if not (direction == BACKWARD or p.hf_hide):
                E = E * _hf_vec(p.carrier, p.nu, t, p.pcos, p.w_list)
            if direction == FORWARD and p.task_type == "intuitive_control":
                for k in range(1, p.impulses_number):
                    E = E + _env_vec(p.envelope, p.E0, t, p.t0 + k * p.delay, p.sigma)
            return list(E)
        out = []
        for t in step_times(p, direction):
            if direction == BACKWARD or p.hf_hide:
                E = p.laser_field(p.E0, t, p.t0, p.sigma)
            else:
                E = p.laser_field(p.E0, t, p.t0, p.sigma) * p.laser_field_hf(np.float64(1.0), t, p.pcos, p.w_list)
            if direction == FORWARD and p.task_type == "intuitive_control":
                for k in range(1, p.impulses_number):
                    E += p.laser_field(p.E0, t, p.t0 + k * p.delay, p.sigma)
            out.append(E)
        return out

    def propagate(self, direction=FORWARD, record=False, chunk=None, on_chunk=None):
        """One prescribed-field sweep of every run from psi0 (forward) / psif (backward)."""
        pl = self.plan
        pl.snapshot_load(0 if direction == FORWARD else 1)
        pl.set_E_base(self._pad([self._prescribed_field(p, direction) for p in self.pbs]))
        slot = self._poly_dir[direction]
        if record:
            pl.record_state(0, 0, 0)
        chunk = chunk or self.nt_max
        l = 1
        while l <= self.nt_max:
            n = min(chunk, self.nt_max - l + 1)
            pl.sweep(slot, eng.FIELD_PRESCRIBED, l, n, True, -1, 0 if record else -1)
            l += n
            if on_chunk:
                on_chunk(l - 1)
        return pl.get_state()

    def _finish_forward(self, it, E0_start):
        pl = self.plan
        E = pl.get_fields()
        if self.pbs[0].task_type == "local_control_projection":
            self.freq_mult = np.ascontiguousarray(E.imag)      # the multiplier rides in the imaginary slot
            self.freq_mult[:, 0] = 1.0
            E = E.real.astype(np.complex128)
        gc = pl.goal_overlap(1)
        rows = []
        exact = not self._vectorised()
        for b, p in enumerate(self.pbs):
            Et = E[b, :p.nt + 1].copy()
            Et[0] = E0_start[b]
            E_int = field_integral(Et, p.t_step, exact)
            F = _F_value(p.F_type, gc[b], p.nb)
            hl = p.h_lambda if p.task_type == "optimal_control_krotov" else np.float64(0.0)
            J = F.real - hl * hl * E_int.real
            self.goal_close_scal[b] += gc[b].sum()
            E_list = _shown_fields(Et)
            rows.append(IterRow(it, float(abs(self.goal_close_scal[b])), float(F.real), float(E_int.real), float(J), E_list))
        return rows, gc

    def run_prescribed(self):
        """single_pot / filtering / trans_wo_control / intuitive_control (fitter.py:425-438)."""
        self.goal_close_scal = np.zeros(self.B, np.complex128)
        pl = self.plan
        pl.snapshot_load(0)
        pl.set_E_base(self._pad([self._prescribed_field(p, FORWARD) for p in self.pbs]))
        self._sweep("iter_0f", FORWARD, (0, 1), 0, eng.FIELD_PRESCRIBED, 1, self.nt_max, -1, -1)
        E0s = [self._prescribed_field(p, FORWARD)[0] for p in self.pbs]
        rows, _ = self._finish_forward(0, E0s)
        return [[r] for r in rows]

    def _local_population_field(self, p):
        """E_l of local_control_population: Euler steps of the field ODE towards the envelope
        (fitter.py:673-692).  The value patched in by do_the_thing (fitter.py:592-595) is overwritten by the
        envelope at the next step (fitter.py:653-658) before anyone reads it, so the field does not depend on
        the state and is integrated here, with the reference's scalar expressions."""
        adt = abs(p.t_step)
        E_vel = np.float64(0.0)
        E = np.float64(0.0)
        out = []
        for t in step_times(p, FORWARD):
            if p.hf_hide:
                patched = p.laser_field(p.E0, t, p.t0, p.sigma)
            else:
                patched = p.laser_field(p.E0, t, p.t0, p.sigma) * p.laser_field_hf(np.float64(1.0), t, p.pcos, p.w_list)
            if E == 0.0:
                E = patched
            if patched <= 0:
                raise RuntimeError("E_patched has to be positive")
            if E <= 0:
                raise RuntimeError("E has to be positive")
            first = E - patched
            second = E_vel * math.pow(patched / E, p.pow_)
            E_acc = -p.k_E * first - p.lamb * second
            E_vel += E_acc * adt
            E = E + E_vel * adt
            out.append(E)
        return out

    def run_local_population(self):
        """local_control_population (fitter.py:425-438 with the field of fitter.py:673-692)."""
        self.goal_close_scal = np.zeros(self.B, np.complex128)
        pl = self.plan
        fields = [self._local_population_field(p) for p in self.pbs]
        pl.snapshot_load(0)
        pl.set_E_base(self._pad(fields))
        self._sweep("iter_0f", FORWARD, (0, 1), 0, eng.FIELD_PRESCRIBED, 1, self.nt_max, -1, -1)
        rows, _ = self._finish_forward(0, [f[0] for f in fields])
        return [[r] for r in rows]

    _LOCAL_ERRORS = {1: "freq_mult_patched has to be positive or zero",
                     2: "Frequency multiplier has to be positive or zero"}

    def run_local_projection(self):
        """local_control_projection: prescribed envelope, carrier-frequency multiplier fed back from
        <psi_e|psi_g> and <psi_e|v_g - v_e|psi_g> of the previous step (fitter.py:598-623, 814-836).  The ODE of
        the multiplier, the energy window and the Newton table of every step are evaluated inside the sweep
        kernel; ``self.freq_mult[run][l]`` holds the multipliers afterwards."""
        p0 = self.pbs[0]
        if not p0.hf_hide or getattr(p0, "carrier", "exp") != "exp":
            raise NotImplementedError("local_control_projection runs in the rotating frame (hf_hide) with the 'exp' "
                                      "carrier; without hf_hide the reference never updates the multiplier")
        self.goal_close_scal = np.zeros(self.B, np.complex128)
        pl = self.plan
        params = np.zeros((self.B, 12))
        for b, p in enumerate(self.pbs):
            kmax = abs(p.akx2[int(p.np_ / 2 - 1)])
            params[b] = [p.v[0][0] + kmax, p.vmin[0], p.v[1][0] + kmax, p.vmin[1], p.E0, p.nu_L,
                         getattr(p, "nu", p.nu_L), p.k_E, p.lamb, p.pow_, p.epsilon, p.t_step]
        xp = math_base.newton_table(p0.nch, 1.0)[0]
        pl.set_local_control(params, self._pad([step_times(p, FORWARD) for p in self.pbs], np.float64), xp)
        pl.snapshot_load(0)
        fields = [self._prescribed_field(p, FORWARD) for p in self.pbs]
        pl.set_E_base(self._pad(fields))
        self._sweep("iter_0f", FORWARD, (0, 1), 0, eng.FIELD_LOCAL_PROJ, 1, self.nt_max, -1, -1)
        self.local_state = pl.get_local_state()
        self.errors = [self._LOCAL_ERRORS.get(int(c), "") for c in self.local_state[:, 4]]
        rows, _ = self._finish_forward(0, [f[0] for f in fields])
        return [[r] for r in rows]

    def prewarm(self):
        """Host tables the control loop needs before its first sweep (the padded prescribed field of iteration 0, the
        carrier at T), computed ahead of time: newcheb.run_variants calls this on its set-up thread, beside the sweeps
        of the previous sub-batch, so that the GPU does not idle while they are evaluated."""
        if getattr(self.pbs[0], "task_type", "") == "optimal_control_krotov":
            self._prewarm_krotov()
        return self

    def _prewarm_krotov(self):
        self._E_presc_fwd = self._pad([self._prescribed_field(p, FORWARD) for p in self.pbs])
        self._sqrt_hf_T = np.array([np.complex128(cmath.sqrt(p.laser_field_hf(np.float64(1.0), p.T, p.pcos, p.w_list)))
                                    for p in self.pbs])

    def run_krotov(self, iter_max=None):
        """optimal_control_krotov for every run (fitter.py:441-489, 537-569).  All runs iterate in
        lock step; a run that has converged keeps its rows and is simply not reported further."""
        pl = self.plan
        hist: List[List[IterRow]] = [[] for _ in self.pbs]
        done = np.zeros(self.B, bool)
        F_mid_1 = np.zeros(self.B)
        self.errors = [""] * self.B
        if getattr(self, "_sqrt_hf_T", None) is None:
            self._prewarm_krotov()
        sqrt_hf_T = self._sqrt_hf_T
        it = 0
        E_start = [self._prescribed_field(p, FORWARD)[0] for p in self.pbs]
        while True:
            self.goal_close_scal = np.zeros(self.B, np.complex128)
            pl.snapshot_load(0)
            pl.record_state(0, 0, 0)                      # psi_tlist[0] = psi_init (fitter.py:227-229)
            if it == 0:
                pl.set_E_base(self._E_presc_fwd)
                self._sweep("iter_0f", FORWARD, (0, 1), 0, eng.FIELD_PRESCRIBED, 1, self.nt_max, -1, 0)
                E0s = E_start
            else:
                self._sweep("iter_%df" % it, FORWARD, (0, 1), 0, eng.FIELD_KROTOV_FWD, 0, self.nt_max + 1, 1, 0)
                E0s = None
            # For it >= 1 every stopping decision below is independent of this sweep's results (the reference tests
            # its loop condition on a zeroed goal_close_vec): if no run goes on, the sweep just launched is the last
            # GPU work of this batch -- a batch entry point may start the next batch's sweeps behind it right now
            if it >= 1 and self.observer is None and self.on_last_sweep is not None:
                limits = [p.iter_max if iter_max is None else iter_max for p in self.pbs]
                if all(done[b] or (lim != -1 and it >= lim) for b, lim in enumerate(limits)):
                    self.on_last_sweep()
            if E0s is None:
                E0s = pl.get_fields()[:, 0]
            rows, gc = self._finish_forward(it, E0s)
            self.last_gc = gc
            stop_all = True
            for b, p in enumerate(self.pbs):
                if done[b]:
                    continue
                limit = p.iter_max if iter_max is None else iter_max
                if it == 0 and abs(rows[b].Fsm - p.F_goal) <= p.epsilon:
                    # goal met by the initial guess: the reference leaves the iteration before reporting it and
                    # before the backward sweep (fitter.py:472-476), so the run has no row at all
                    done[b] = True
                    continue
                hist[b].append(rows[b])
                # The loop condition of time_propagation (fitter.py:545-546) is tested after the BACKWARD sweep of
                # the iteration, which recomputes Fsm from a zeroed goal_close_vec (fitter.py:354-410): it sees
                # F_type(0) = 0, not the forward value, so a Krotov run always uses its whole iteration budget
                # unless |0 - F_goal| <= epsilon.  The divergence guard (fitter.py:550-560) sees the same zeros.
                F_after = 0.0
                if p.q and it >= 1:
                    if it == p.iter_mid_1:
                        F_mid_1[b] = F_after
                    elif it == p.iter_mid_2:
                        lim = -p.nb * p.nb * p.q
                        if F_mid_1[b] > lim or F_after > lim or F_mid_1[b] < F_after:
                            self.errors[b] = "The calculation is probably going to diverge.  Stopping the run..."
                            done[b] = True
                            continue
                if abs(F_after - p.F_goal) <= p.epsilon or (it >= limit and limit != -1):
                    done[b] = True
                else:
                    stop_all = False
            # the reference always runs the backward sweep of an iteration it started, unless the goal was met in
            # iteration 0 (fitter.py:472-476); its results only matter to the next iteration -- and to the per-step
            # tables of "iter_{k}b" when reporters are attached
            if stop_all and (self.observer is None or not any(hist)):
                break
            pl.krotov_terminal(1, sqrt_hf_T, 1)
            pl.snapshot_set(2)                            # chi(T): the "psi0" of the backward solver (fitter.py:247)
            self._sweep("iter_%db" % it, BACKWARD, (2, 0), 1, eng.FIELD_KROTOV_BWD, 1, self.nt_max, 0, 1)
            if stop_all:
                break
            it += 1
        return hist


class NLevelBatch(_BatchBase):
    """N-level (np = 1) runs: unitary-transformation optimal control and prescribed fields."""

    def _upload_static(self):
        uwd = np.array([[float(p.U), float(p.W), float(p.delta)] for p in self.pbs])
        self.plan.set_static(np.stack([p.v for p in self.pbs]), None, uwd, self.dx)

    def _vectorised(self):
        return int(self.nt.sum()) > VECTOR_SETUP_POINTS

    def _field_rows(self, direction, half_step=False):
        rows = []
        vec = self._vectorised()
        for p in self.pbs:
            env, hf = field_tables(p, direction, half_step, vec)
            rows.append(env * hf)
        return rows

    def run_prescribed(self):
        """trans_wo_control on an N-level system, e.g. the pi-pulse test (functional_tests.py:1072-1164)."""
        pl = self.plan
        pl.set_control(1.0, self.nt.astype(np.float64))
        rows = self._field_rows(FORWARD)
        pl.set_E_base(self._pad(rows))
        pl.snapshot_load(0)
        self._sweep("iter_0f", FORWARD, (0, 1), 0, eng.FIELD_PRESCRIBED, 1, self.nt_max, -1, -1)
        E = pl.get_fields()
        gc = pl.goal_overlap(1)
        out = []
        for b, p in enumerate(self.pbs):
            Et = np.array([rows[b][0]] + [E[b, l] for l in range(1, p.nt + 1)])
            E_int = field_integral(Et, p.t_step)
            F = _F_value(p.F_type, gc[b], p.nb)
            E_list = _shown_fields(Et)
            out.append([IterRow(0, float(abs(gc[b].sum())), float(F.real), float(E_int.real), float(F.real), E_list)])
        return out

    def _initial_row(self, b):
        """iteration -1 report (fitter.py:494-535)."""
        p = self.pbs[b]
        gc0 = np.zeros(p.nb, np.complex128)
        sc = np.float64(0.0)
        for v in range(p.nb):
            for n in range(p.nlevs):
                gc0[v] += np.vdot(p.psi0[v][n], p.psif[v][n]) * p.dx
            sc += gc0[v]
        F0 = _F_value(p.F_type, gc0, p.nb)
        if self._vectorised() and getattr(p, "envelope", None):
            t = np.asarray(p.t_list, np.float64)
            El = np.ascontiguousarray((_env_vec(p.envelope, p.E0, t, p.t0, p.sigma) *
                                       _hf_vec(p.carrier, p.nu, t, p.pcos, p.w_list)).real)
            Ei = field_integral(El, p.t_step, exact=False).real
        else:
            El = [(p.laser_field(p.E0, t, p.t0, p.sigma) * p.laser_field_hf(np.float64(1.0), t, p.pcos, p.w_list)).real
                  for t in p.t_list]
            Ei = field_integral(np.array([0.0] + list(El[1:])), p.t_step).real
        h0 = p.h_lambda * RED_PLANCK_H / CM_TO_ERG if p.ntriv == -1 else p.h_lambda
        return IterRow(-1, float(abs(sc)), float(F0.real), float(Ei), float(F0.real - h0 * h0 * Ei), np.array(El))

    def run_unit_transform(self, iter_max=None, keep_fields=True):
        """optimal_control_unit_transform for every run (fitter.py:441-569, 709-778, 799-807).

        keep_fields=True copies every run's field list to the host after every forward sweep (what
        print_iter_point_fitter receives, and what the parity tests compare).  keep_fields=False is the
        throughput mode for big batches whose reports only need the iteration table (plotting_flag
        "tables_iter", input_batch/input_report.json:7): E_int is reduced on the device and the per-iteration
        host work is a handful of numpy array operations, so the loop is bound by the sweeps themselves."""
        pl = self.plan
        B = self.B
        kind = self.pbs[0].F_type
        assert keep_fields or all(p.F_type == kind for p in self.pbs)
        h0_all = np.array([p.h_lambda * RED_PLANCK_H / CM_TO_ERG if p.ntriv == -1 else p.h_lambda for p in self.pbs])
        mode_dyn = np.array([p.h_lambda_mode == "dynamical" for p in self.pbs])
        nb_runs = np.array([float(p.nb) for p in self.pbs])
        t_steps = np.array([float(p.t_step) for p in self.pbs])
        hist: List[List[IterRow]] = [[self._initial_row(b)] for b in range(B)]
        self.errors = [""] * B
        done = np.zeros(B, bool)
        goal_close_abs = np.zeros(B)
        F_mid_1 = np.zeros(B)
        h_lambda = np.ones(B)
        a0 = np.zeros((B, self.pbs[0].nb), np.complex128)
        s_rows = []
        base_rows = []
        vec = self._vectorised()
        for p in self.pbs:
            env, hf = field_tables(p, FORWARD, True, vec)
            sr = env / p.E0
            sr[0] = 0.0
            br = sr * p.E0 * hf                          # s * E0 * hf(t - |dt|/2)   (fitter.py:753-755)
            br[0] = env[0] * hf[0]                       # E_patched at t = 0        (fitter.py:653-658)
            s_rows.append(sr)
            base_rows.append(br)
        pl.set_s_env(self._pad(s_rows))
        it = 0
        # throughput mode bookkeeping as arrays over the batch (rows are materialised after the loop)
        q_all = np.array([float(p.q) for p in self.pbs])
        mid1_all = np.array([p.iter_mid_1 for p in self.pbs])
        mid2_all = np.array([p.iter_mid_2 for p in self.pbs])
        lim_all = np.array([-p.nb * p.nb * float(p.q) for p in self.pbs])
        Fgoal_all = np.array([float(p.F_goal) for p in self.pbs])
        eps_all = np.array([float(p.epsilon) for p in self.pbs])
        itmax_all = np.array([p.iter_max if iter_max is None else iter_max for p in self.pbs])
        records = []
        self.ctx.synchronize()
        t_loop = time.perf_counter()
        self.iter_marks = []                      # host clock at the start of every iteration (tools/bench_nlevel.py)
        while True:
            self.iter_marks.append(time.perf_counter())
            # ---- backward sweep from the goal basis (fitter.py:241-249) ----------------------------
            pl.snapshot_load(1)
            pl.record_state(1, 0)
            if it > 0:                                    # finished runs leave the sweeps (their warps retire at once)
                pl.set_control(h_lambda, self.nt.astype(np.float64), retired=done)
            if it == 0:
                pl.set_control(1.0, self.nt.astype(np.float64))
                pl.set_E_base(self._pad(self._field_rows(BACKWARD)))
                self._sweep("iter_0b", BACKWARD, (1, 0), 1, eng.FIELD_PRESCRIBED, 1, self.nt_max, -1, 1)
            else:
                self._sweep("iter_%db" % it, BACKWARD, (1, 0), 1, eng.FIELD_PRESCRIBED, 1, self.nt_max, -1, 1,
                            reversed_E=True)
            chi_init = pl.get_state()
            # ---- l == 0 of the forward sweep: h_lambda and a0 (fitter.py:712-737) -------------------
            if keep_fields:
                for b, p in enumerate(self.pbs):
                    h0 = p.h_lambda * RED_PLANCK_H / CM_TO_ERG if p.ntriv == -1 else p.h_lambda
                    if p.h_lambda_mode == "dynamical" and goal_close_abs[b]:
                        h_lambda[b] = h0 * p.nb / math.sqrt(goal_close_abs[b])
                    else:
                        h_lambda[b] = h0
                    a0[b] = _a_value(p.F_type, p.psi0, chi_init[b], p.dx)
            else:
                # the same formulas over the whole batch at once (numpy); summation order differs from the
                # reference's nested loops in the last bit
                dyn = mode_dyn & (goal_close_abs != 0.0)
                h_lambda = np.where(dyn, h0_all * nb_runs / np.sqrt(np.where(dyn, goal_close_abs, 1.0)), h0_all)
                a0 = a_batch(kind, self.psi0, chi_init, self.dx)
            pl.set_control(h_lambda, self.nt.astype(np.float64), a0, retired=done)
            if it == 0:
                pl.set_E_base(self._pad(base_rows))
            pl.snapshot_load(0)
            self._sweep("iter_%df" % it, FORWARD, (0, 1), 0, eng.FIELD_UT_FWD, 0, self.nt_max + 1, 1, -1,
                        write_E_base=True)
            gc = pl.goal_overlap(1)
            if keep_fields:
                E = pl.get_fields()
            else:
                # throughput mode: E_int reduced on the device, functional over the whole batch at once
                E_int_all = pl.field_integral() * (t_steps * 1e15)
                tau = gc.sum(axis=1)
                F_all = F_batch(kind, gc)
            if not keep_fields:
                active = ~done
                J_all = F_all - h_lambda * h_lambda * E_int_all
                goal_close_abs = np.where(active, np.abs(tau), goal_close_abs)
                records.append((it, active, goal_close_abs.copy(), F_all, E_int_all, J_all))
                guard = active & (q_all != 0.0) & (it >= 1)
                at1 = guard & (it == mid1_all)
                F_mid_1 = np.where(at1, F_all, F_mid_1)
                at2 = guard & ~at1 & (it == mid2_all)
                diverged = at2 & ((F_mid_1 > lim_all) | (F_all > lim_all) | (F_mid_1 < F_all))     # fitter.py:550-560
                for b in np.nonzero(diverged)[0]:
                    self.errors[b] = "The calculation is probably going to diverge.  Stopping the run..."
                reached = (np.abs(F_all - Fgoal_all) <= eps_all) | ((it >= itmax_all) & (itmax_all != -1))
                done = done | diverged | (active & reached)
                if done.all():
                    break
                it += 1
                continue
            stop_all = True
            for b, p in enumerate(self.pbs):
                if done[b]:
                    continue
                if True:
                    Et = E[b, :p.nt + 1]
                    E_int = field_integral(Et, p.t_step)
                    F = _F_value(p.F_type, gc[b], p.nb)
                    E_list = _shown_fields(Et)
                J = F.real - h_lambda[b] * h_lambda[b] * E_int.real
                goal_close_abs[b] = abs(gc[b].sum())
                hist[b].append(IterRow(it, float(goal_close_abs[b]), float(F.real), float(E_int.real), float(J), E_list))
                if p.q and it >= 1:                       # divergence guard (fitter.py:550-560)
                    if it == p.iter_mid_1:
                        F_mid_1[b] = F.real
                    elif it == p.iter_mid_2:
                        lim = -p.nb * p.nb * p.q
                        if F_mid_1[b] > lim or F.real > lim or F_mid_1[b] < F.real:
                            self.errors[b] = "The calculation is probably going to diverge.  Stopping the run..."
                            done[b] = True
                            continue
                limit = p.iter_max if iter_max is None else iter_max
                if abs(F.real - p.F_goal) <= p.epsilon or (it >= limit and limit != -1):
                    done[b] = True
                else:
                    stop_all = False
            if stop_all:
                break
            it += 1
        self.ctx.synchronize()
        self.iter_marks.append(time.perf_counter())
        self.loop_wall_s = time.perf_counter() - t_loop       # iteration loop only (set-up excluded)
        for k, active, gca, F_all, E_int_all, J_all in records:
            for b in np.nonzero(active)[0]:
                hist[b].append(IterRow(k, float(gca[b]), float(F_all[b]), float(E_int_all[b]), float(J_all[b]), None))
        return hist
