"""GPU-backed counterparts of the reference's numerical kernels (phys_base.py:11-288), same names
and argument meaning, so code written against the reference's module can switch imports:

    prop_cpu(psi, hamil2D, t_sc, nch, np, emin, emax, E, eL, E_full, orig)   one polynomial step
    hamil_apply(hamil2D, orig, psi, E, eL, E_full)                           what hamil2D(...) does
    diff_cpu(psi, akx2, np)                                                  kinetic mapping of one level
    exp_vals_calc / exp_sigma_vals_calc                                      expectation values

The names keep the reference's "_cpu" suffix because they are drop-ins; every one of them runs on
the GPU through the C ABI (a per-descriptor plan with batch = 1) and raises if that is impossible.
This is the slow, call-by-call compatibility path; the batched sweeps live in control.py.
"""
from __future__ import annotations

import cmath
import threading

import numpy

from . import engine as _eng
from . import math_base
from .phys_consts import CM_TO_ERG, DALT_TO_AU, HART_TO_CM, HZ_TO_CM, RED_PLANCK_H
from .psi_basis import Psi

hart_to_cm = HART_TO_CM
cm_to_erg = CM_TO_ERG
dalt_to_au = DALT_TO_AU
Red_Planck_h = RED_PLANCK_H
Hz_to_cm = HZ_TO_CM


def func(z, t):
    """exp(-i z t): the function the Newton polynomial interpolates (phys_base.py:116-126)."""
    return cmath.exp(-1j * z * t)


class ExpectationValues:
    def __init__(self, x, x2, p, p2):
        self.x, self.x2, self.p, self.p2 = x, x2, p, p2


class SigmaExpectationValues:
    def __init__(self, x, y, z):
        self.x, self.y, self.z = x, y, z


# ----------------------------------------------------------------------------- plan cache
# The reference farms its jobs over a thread pool (newcheb.py:656-659), so this layer must be callable from several
# host threads at once.  Every thread gets its OWN context (device binding + stream) and its own session cache:
# nothing mutable is shared between threads, calls of different threads run concurrently on different streams
# (ctypes releases the GIL), and the sequences upload -> step -> download of two threads can never interleave on
# one plan.  The contexts of a thread are released when the thread ends.
_TLS = threading.local()


def _tls():
    if not hasattr(_TLS, "ctx"):
        _TLS.ctx, _TLS.sessions = {}, {}
    return _TLS


def default_context(device: int = 0) -> _eng.Context:
    ctxs = _tls().ctx
    if device not in ctxs:
        ctxs[device] = _eng.Context(device)
    return ctxs[device]


class Session:
    """A batch-of-one plan bound to one Hamiltonian descriptor (Hamil2D*)."""

    def __init__(self, hamil2D, nlevs, nch, hf_hide=False, dx=1.0, x=None, device=0):
        self.h = hamil2D
        self.nlevs, self.nch, self.np = int(nlevs), int(nch), int(hamil2D._np)
        self.ctx = default_context(device)
        self.plan = _eng.Plan(self.ctx, hamil2D.kind, self.np, self.nlevs, 1, self.nch, 1, 1, hf_hide=hf_hide)
        grid = hamil2D.kind == _eng.HAMIL_NONTRIV
        self.plan.set_static(hamil2D.potentials(), hamil2D.kinetic() if grid else None,
                             None if grid or hamil2D.kind == _eng.HAMIL_TWO_LEVELS
                             else numpy.array([float(hamil2D._U), float(hamil2D._W), float(hamil2D._delta)]), float(dx))
        self.dx = float(dx)
        self._poly_key = None
        if x is not None:
            akx = None
            if grid:
                akx = math_base.initak(self.np, dx, 1, hamil2D._ntriv) * (hart_to_cm / (-1j) / dalt_to_au)
                akx = numpy.ascontiguousarray(akx.real)
            self.plan.set_instr_grid(numpy.asarray(x, numpy.float64), akx)

    def set_poly(self, t_sc, emin, emax, eL):
        key = (float(t_sc), float(emin), float(emax), float(eL))
        if key != self._poly_key:
            xp, dv = math_base.newton_table(self.nch, t_sc)
            self.plan.set_poly(0, xp, dv, emin, emax, t_sc, eL)
            self._poly_key = key

    def upload(self, psi: Psi):
        self.plan.set_state(psi.as_array()[None, None])

    def download(self) -> Psi:
        return Psi.from_array(self.plan.get_state()[0, 0])


def session_for(hamil2D, nlevs, nch, dx=1.0) -> Session:
    sessions = _tls().sessions
    key = (id(hamil2D), int(nlevs), int(nch), float(dx))
    s = sessions.get(key)
    if s is None or s.h is not hamil2D:
        if len(sessions) > 32:
            for old in list(sessions.values()):
                old.plan.close()
            sessions.clear()
        s = Session(hamil2D, nlevs, nch, dx=dx)
        sessions[key] = s
    return s


# ----------------------------------------------------------------------------- drop-ins
def prop_cpu(psi: Psi, hamil2D, t_sc, nch, np, emin, emax, E, eL, E_full, orig):
    """psi <- P(Hn) psi: Newton interpolation of exp(-i Hn t_sc) at nch Chebyshev nodes, on the GPU.
    Mutates and returns ``psi`` like the reference (phys_base.py:129-178)."""
    for k in range(len(psi.f)):
        assert psi.f[k].size == np
    if orig:
        raise NotImplementedError("prop_cpu(orig=True) is never used by the reference's stepper (propagation.py:420,445)")
    s = session_for(hamil2D, len(psi.f), nch)
    s.set_poly(t_sc, emin, emax, eL)
    s.upload(psi)
    grid = hamil2D.kind == _eng.HAMIL_NONTRIV
    s.plan.prop_step(0, E if grid or hamil2D.kind in (_eng.HAMIL_BH_X, _eng.HAMIL_BH_Z) else E_full)
    out = s.plan.get_state()[0, 0]
    for k in range(len(psi.f)):
        psi.f[k][...] = out[k]
    return psi


def hamil_apply(hamil2D, orig, psi: Psi, E, eL, E_full) -> Psi:
    """phi = H psi, the body of Hamil2D*.__call__ (hamil_2d.py:33-187), on the GPU."""
    s = session_for(hamil2D, len(psi.f), 2)
    s.upload(psi)
    kind = hamil2D.kind
    if kind == _eng.HAMIL_NONTRIV:
        field = E_full if orig else E
    elif kind in (_eng.HAMIL_BH_X, _eng.HAMIL_BH_Z):
        field = E
    else:
        field = E_full
    return Psi.from_array(s.plan.apply_hamil(0, bool(orig), field, float(numpy.real(eL)))[0, 0])


def diff_cpu(psi, akx2, np):
    """ifft(fft(psi) * akx2) for one level (phys_base.py:18-36), through the kinetic application of a
    throw-away two-level grid plan."""
    from .hamil_2d import Hamil2DNonTrivial
    assert psi.size == np and akx2.size == np
    zeros = numpy.zeros(np)
    h = Hamil2DNonTrivial([(0.0, zeros), (0.0, zeros)], numpy.asarray(akx2), np, 0.0, 0.0, 1.0, 1)
    s = Session(h, 2, 2)
    s.plan.set_state(numpy.stack([psi, numpy.zeros(np, numpy.complex128)])[None, None])
    out = s.plan.apply_hamil(1, False, 0.0, 0.0)[0, 0, 0].copy()
    s.plan.close()
    return out
