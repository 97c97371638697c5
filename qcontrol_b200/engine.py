"""numpy-facing wrapper of the C ABI: contexts, plans and batched sweeps.

This is the layer the reference-shaped classes (``propagation.PropagationSolver``,
``fitter.FittingSolver``, ``phys_base.prop_cpu``) are written on.  It only moves
data and launches kernels; there is no arithmetic fallback here.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional

import numpy as np

from . import _lib
from ._lib import PlanDesc, SweepArgs, check, lib

HAMIL_NONTRIV, HAMIL_TWO_LEVELS, HAMIL_BH_X, HAMIL_BH_Z, HAMIL_QUBIT = range(5)
FIELD_PRESCRIBED, FIELD_KROTOV_FWD, FIELD_KROTOV_BWD, FIELD_UT_FWD, FIELD_LOCAL_PROJ = range(5)

_DP = C.POINTER(C.c_double)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        assert a.shape == tuple(shape), (a.shape, shape)
    return a


def _c128(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.complex128)
    if shape is not None:
        assert a.shape == tuple(shape), (a.shape, shape)
    return a


def _ptr(a):
    return a.ctypes.data_as(_DP) if a is not None else None


class Context:
    """One CUDA device + one stream (qc_create)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._h = C.c_void_p()
        check(lib.qc_create(int(device), C.c_void_p(stream) if stream else None, C.byref(self._h)))
        self.device = int(device)
        self._plans = weakref.WeakSet()

    def close(self):
        for pl in list(self._plans):      # plans hold a pointer to the context: they go first
            pl.close()
        if self._h:
            lib.qc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(lib.qc_synchronize(self._h))

    def timer_start(self):
        check(lib.qc_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        check(lib.qc_timer_stop(self._h, C.byref(ms)))
        return ms.value

    @property
    def launches(self) -> int:
        return int(lib.qc_launch_count(self._h))

    def mem_info(self):
        """(free, total) bytes of the device's memory."""
        f, t = C.c_int64(), C.c_int64()
        check(lib.qc_mem_info(self._h, C.byref(f), C.byref(t)))
        return int(f.value), int(t.value)

    @property
    def sm_count(self) -> int:
        n = C.c_int32()
        check(lib.qc_sm_count(self._h, C.byref(n)))
        return int(n.value)

    def pool_reserve(self, nbytes: int):
        """Reserve an arena of device memory for the plans of this device (0 gives it back), see qc_pool_reserve."""
        check(lib.qc_pool_reserve(self._h, int(nbytes)))

    def fp64_peak_tflops(self) -> float:
        tf = C.c_double()
        check(lib.qc_measure_fp64_peak(self._h, C.byref(tf)))
        return tf.value


def plan_footprint(hamil, np_, nlevs, nb, nch, batch, nt_max, hf_hide=False, store_traj=False) -> int:
    """Upper bound of the device bytes a plan of this shape allocates over its life (qc_plan_footprint)."""
    d = PlanDesc(int(hamil), int(np_), int(nlevs), int(nb), int(nch), int(batch), int(nt_max), int(bool(hf_hide)),
                 int(bool(store_traj)), 0)
    out = C.c_int64()
    check(lib.qc_plan_footprint(C.byref(d), C.byref(out)))
    return int(out.value)


class Plan:
    """Device-resident batch of runs of one shape (qc_plan_create)."""

    def __init__(self, ctx: Context, hamil: int, np_: int, nlevs: int, nb: int, nch: int, batch: int,
                 nt_max: int, hf_hide: bool = False, store_traj: bool = False):
        self.ctx = ctx
        self.hamil, self.np, self.nlevs, self.nb, self.nch = int(hamil), int(np_), int(nlevs), int(nb), int(nch)
        self.batch, self.nt_max = int(batch), int(nt_max)
        self.hf_hide, self.store_traj = bool(hf_hide), bool(store_traj)
        d = PlanDesc(self.hamil, self.np, self.nlevs, self.nb, self.nch, self.batch, self.nt_max,
                     int(self.hf_hide), int(self.store_traj), 0)
        self._h = C.c_void_p()
        check(lib.qc_plan_create(ctx._h, C.byref(d), C.byref(self._h)))
        ctx._plans.add(self)
        self.state_shape = (self.batch, self.nb, self.nlevs, self.np)

    def close(self):
        if self._h:
            lib.qc_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- uploads ---------------------------------------------------------------
    def set_static(self, v, akx2_re=None, uwd=None, dx=1.0):
        v = _f64(v)
        vb = v.ndim == 3
        _f64(v, ((self.batch,) if vb else ()) + (self.nlevs, self.np))
        k = _f64(akx2_re, (self.np,)) if akx2_re is not None else None
        u = None
        ub = 0
        if uwd is not None:
            u = _f64(uwd)
            ub = int(u.ndim == 2)
            _f64(u, ((self.batch,) if ub else ()) + (3,))
        dxa = _f64(np.atleast_1d(dx))
        db = int(dxa.shape[0] == self.batch and self.batch > 1)
        assert dxa.shape[0] in (1, self.batch)
        check(lib.qc_set_static(self._h, _ptr(v), int(vb), _ptr(k), _ptr(u), ub, _ptr(dxa), db))

    def set_poly(self, slot, xp, dv, emin, emax, t_sc, eL=0.0):
        xp = _f64(xp, (self.nch,))
        dv = _c128(dv)
        batched = dv.ndim == 2
        _c128(dv, ((self.batch,) if batched else ()) + (self.nch,))
        rows = self.batch if batched else 1
        sc = np.empty((rows, 4), np.float64)
        sc[:, 0], sc[:, 1], sc[:, 2], sc[:, 3] = emin, emax, t_sc, eL
        check(lib.qc_set_poly(self._h, int(slot), _ptr(xp), _ptr(dv.view(np.float64)), _ptr(sc), int(batched)))
        self._poly_batched = getattr(self, "_poly_batched", {})
        self._poly_batched[int(slot)] = bool(batched)

    def set_poly_orders(self, slot, orders=None):
        """Opt-in truncation of the interpolation series of ``slot`` to orders[run] terms (None: the full series)."""
        if orders is None:
            check(lib.qc_set_poly_orders(self._h, int(slot), None, 0))
            return
        batched = getattr(self, "_poly_batched", {}).get(int(slot), False)      # same row layout as set_poly's dv
        o = np.ascontiguousarray(np.atleast_1d(orders), dtype=np.int32)
        assert o.shape[0] == (self.batch if batched else 1), (o.shape, batched, self.batch)
        check(lib.qc_set_poly_orders(self._h, int(slot), o.ctypes.data_as(C.POINTER(C.c_int32)), int(batched)))

    def set_state(self, psi):
        psi = _c128(psi, self.state_shape)
        check(lib.qc_set_state(self._h, _ptr(psi.view(np.float64))))

    def get_state(self):
        out = np.empty(self.state_shape, np.complex128)
        check(lib.qc_get_state(self._h, _ptr(out.view(np.float64))))
        return out

    def _set_scalars(self, which, vals):
        vals = _c128(vals)
        batched = vals.ndim == 2
        _c128(vals, ((self.batch,) if batched else ()) + (self.nt_max + 1,))
        check(lib.qc_set_step_scalars(self._h, which, _ptr(vals.view(np.float64)), int(batched)))

    def set_E_base(self, E):
        self._set_scalars(0, E)

    def set_exp_L(self, e, slot=0):
        self._set_scalars(1 if slot == 0 else 3, e)

    def set_s_env(self, s):
        self._set_scalars(2, s)

    def set_control(self, h_lambda, nt, a0=None, retired=None):
        """Per-run control scalars (h_lambda, nt) and, for N-level plans, the 'a' coefficients; ``retired[batch]``
        marks runs that take no further part in sweeps (converged / stopped runs of a batch keep their slot)."""
        h = np.atleast_1d(np.asarray(h_lambda, np.float64))
        n = np.atleast_1d(np.asarray(nt, np.float64))
        rows = max(h.shape[0], n.shape[0])
        if retired is not None:
            rows = self.batch
        assert rows in (1, self.batch)
        ctl = np.zeros((rows, 4), np.float64)
        ctl[:, 0], ctl[:, 1] = h, n
        if retired is not None:
            ctl[:, 2] = np.asarray(retired, np.float64)
        a = _c128(a0, (self.batch, self.nb)) if a0 is not None else None
        check(lib.qc_set_control(self._h, _ptr(ctl), int(rows == self.batch and self.batch > 1),
                                 _ptr(a.view(np.float64)) if a is not None else None))

    # -- compute ---------------------------------------------------------------
    @staticmethod
    def _args(slot, mode, l_first, nsteps, renorm, traj_read, traj_write, reversed_E, write_E_base):
        return SweepArgs(int(slot), int(mode), int(l_first), int(nsteps), int(bool(renorm)), int(traj_read),
                         int(traj_write), int(bool(reversed_E)), int(bool(write_E_base)), (C.c_int32 * 3)(0, 0, 0))

    def sweep(self, slot=0, mode=FIELD_PRESCRIBED, l_first=1, nsteps=1, renorm=True, traj_read=-1, traj_write=-1,
              reversed_E=False, write_E_base=False):
        a = self._args(slot, mode, l_first, nsteps, renorm, traj_read, traj_write, reversed_E, write_E_base)
        check(lib.qc_sweep(self._h, C.byref(a)))

    def prop_step(self, slot, E):
        E = _c128(np.broadcast_to(np.asarray(E, np.complex128), (self.batch,)))
        check(lib.qc_prop_step(self._h, int(slot), _ptr(E.view(np.float64))))

    def propagate_host(self, psi_in, psi_out, slot=0, l_first=1, nsteps=1, renorm=True):
        a = self._args(slot, FIELD_PRESCRIBED, l_first, nsteps, renorm, -1, -1, False, False)
        assert psi_in.dtype == np.complex128 and psi_in.flags.c_contiguous and psi_in.shape == self.state_shape
        assert psi_out.dtype == np.complex128 and psi_out.flags.c_contiguous and psi_out.shape == self.state_shape
        check(lib.qc_propagate_host(self._h, C.byref(a), _ptr(psi_in.view(np.float64)), _ptr(psi_out.view(np.float64))))

    def set_local_control(self, params, t_tab, xp):
        """Constants of the fed-back multiplier ODE, params[batch][12] = extr(4), E0, nu_L, nu, k_E, lamb, pow,
        epsilon, dt; step times t_tab[batch][nt_max+1]; nodes xp[nch].  Resets the feedback state."""
        params = np.ascontiguousarray(params, np.float64)
        t_tab = np.ascontiguousarray(t_tab, np.float64)
        xp = np.ascontiguousarray(xp, np.float64)
        assert params.shape == (self.batch, 12) and t_tab.shape == (self.batch, self.nt_max + 1) and xp.shape == (self.nch,)
        check(lib.qc_set_local_control(self._h, _ptr(params), _ptr(t_tab), _ptr(xp)))

    def get_local_state(self):
        """[batch][11]: freq_mult, velocity, patched value, dAdt_happy, error code, psigc_psie, psigc_dv_psie,
        dA/dt, error step."""
        out = np.empty((self.batch, 11), np.float64)
        check(lib.qc_get_local_state(self._h, _ptr(out)))
        return out

    def set_local_state(self, state):
        """state[batch][4] = freq_mult, freq_mult_vel, freq_mult_patched, dAdt_happy (restart from a stored step)."""
        state = np.ascontiguousarray(state, np.float64)
        assert state.shape == (self.batch, 4)
        check(lib.qc_set_local_state(self._h, _ptr(state)))

    def get_fields(self):
        out = np.empty((self.batch, self.nt_max + 1), np.complex128)
        check(lib.qc_get_fields(self._h, _ptr(out.view(np.float64))))
        return out

    def field_integral(self):
        """sum_{l>=1} |E_l|^2 per run of the last sweep, reduced on the device."""
        out = np.empty(self.batch, np.float64)
        check(lib.qc_field_integral(self._h, _ptr(out)))
        return out

    def get_norms(self):
        out = np.empty((self.batch, self.nb, self.nlevs), np.float64)
        check(lib.qc_get_norms(self._h, _ptr(out)))
        return out

    def snapshot_set(self, slot, psi=None):
        """Keep a full state resident in HBM (slot 0..3); psi=None snapshots the current state."""
        if psi is None:
            check(lib.qc_snapshot_set(self._h, int(slot), None))
        else:
            psi = _c128(psi, self.state_shape)
            check(lib.qc_snapshot_set(self._h, int(slot), _ptr(psi.view(np.float64))))

    def snapshot_load(self, slot):
        check(lib.qc_snapshot_load(self._h, int(slot)))

    def goal_overlap(self, goal_slot, fetch=True):
        if not fetch:
            check(lib.qc_goal_overlap(self._h, int(goal_slot), None))
            return None
        out = np.empty((self.batch, self.nb), np.complex128)
        check(lib.qc_goal_overlap(self._h, int(goal_slot), _ptr(out.view(np.float64))))
        return out

    def krotov_terminal(self, goal_slot, sqrt_hf_T, traj_write, gc=None):
        g = _c128(gc, (self.batch, self.nb)) if gc is not None else None
        hf = _c128(np.atleast_1d(sqrt_hf_T))
        assert hf.shape[0] in (1, self.batch)
        check(lib.qc_krotov_terminal(self._h, int(goal_slot), _ptr(g.view(np.float64)) if g is not None else None,
                                     _ptr(hf.view(np.float64)), int(hf.shape[0] == self.batch and self.batch > 1),
                                     int(traj_write)))

    def get_traj(self, buffer, index):
        nrec = 1 if self.hamil == HAMIL_NONTRIV else self.nb * self.nlevs
        out = np.empty((self.batch, nrec, self.np), np.complex128)
        check(lib.qc_get_traj(self._h, int(buffer), int(index), _ptr(out.view(np.float64))))
        return out

    def record_state(self, buffer, index, level=0):
        check(lib.qc_record_state(self._h, int(buffer), int(index), int(level)))

    # -- bare operator applications / instrumentation ------------------------------------------
    def apply_hamil(self, which=0, orig=False, E=0.0, eL=0.0):
        """phi = Op psi of the resident state (state unchanged): which 0 = H (orig or shifted form),
        1 = kinetic operator, 2 = momentum operator."""
        Ea = _c128(np.broadcast_to(np.asarray(E, np.complex128), (self.batch,)))
        out = np.empty(self.state_shape, np.complex128)
        check(lib.qc_apply_hamil(self._h, int(which), int(bool(orig)), _ptr(Ea.view(np.float64)), float(eL),
                                 _ptr(out.view(np.float64))))
        return out

    def set_instr_grid(self, x, akx_re=None):
        x = _f64(x, (self.np,))
        k = _f64(akx_re, (self.np,)) if akx_re is not None else None
        check(lib.qc_set_instr_grid(self._h, _ptr(x), _ptr(k)))

    def instrument(self, E_full=0.0, two_m=1.0, psi0_slot=0, psif_slot=1):
        """Device-side instrumentation table [batch][nb][nlevs][20] (see include/qcontrol_b200.h)."""
        Ea = _c128(np.broadcast_to(np.asarray(E_full, np.complex128), (self.batch,)))
        out = np.empty(self.state_shape[:3] + (20,), np.float64)
        check(lib.qc_instrument(self._h, _ptr(Ea.view(np.float64)), float(two_m), int(psi0_slot), int(psif_slot),
                                _ptr(out)))
        return out


class Plan2D:
    """Device-resident batch of one-level wavefunctions on an n x n grid (qc_plan2d_*, kernel C): the synthetic
    "hamil_2d 512 x 512" extension.  psi[batch][n][n], row-major (y contiguous)."""

    def __init__(self, ctx: Context, n: int, nch: int, batch: int):
        self.ctx = ctx
        self.n, self.nch, self.batch = int(n), int(nch), int(batch)
        self._h = C.c_void_p()
        check(lib.qc_plan2d_create(ctx._h, self.n, self.nch, self.batch, C.byref(self._h)))
        ctx._plans.add(self)
        self.state_shape = (self.batch, self.n, self.n)

    def close(self):
        if self._h:
            lib.qc_plan2d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_static(self, v, akx2_re, aky2_re, dxdy):
        v = _f64(v)
        vb = v.ndim == 3
        _f64(v, ((self.batch,) if vb else ()) + (self.n, self.n))
        kx, ky = _f64(akx2_re, (self.n,)), _f64(aky2_re, (self.n,))
        check(lib.qc_plan2d_set_static(self._h, _ptr(v), int(vb), _ptr(kx), _ptr(ky), float(dxdy)))

    def set_poly(self, xp, dv, emin, emax, t_sc):
        xp = _f64(xp, (self.nch,))
        dv = _c128(dv, (self.nch,))
        sc = np.array([emin, emax, t_sc], np.float64)
        check(lib.qc_plan2d_set_poly(self._h, _ptr(xp), _ptr(dv.view(np.float64)), _ptr(sc)))

    def set_state(self, psi):
        psi = _c128(psi, self.state_shape)
        check(lib.qc_plan2d_set_state(self._h, _ptr(psi.view(np.float64))))

    def get_state(self):
        out = np.empty(self.state_shape, np.complex128)
        check(lib.qc_plan2d_get_state(self._h, _ptr(out.view(np.float64))))
        return out

    def sweep(self, nsteps=1, renorm=True):
        check(lib.qc_plan2d_sweep(self._h, int(nsteps), int(bool(renorm))))

    def get_norms(self):
        out = np.empty(self.batch, np.float64)
        check(lib.qc_plan2d_get_norms(self._h, _ptr(out)))
        return out
