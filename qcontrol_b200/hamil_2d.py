"""Hamiltonian descriptors with the reference's functor interface (hamil_2d.py:190-223).

``Hamil2D*(v, akx2, np, U, W, delta, ntriv)`` carries the static data of one run; calling it,
``hamil2D(orig, psi, E, eL, E_full) -> Psi``, applies H once ON THE GPU (used by the energy
instrumentation, propagation.py:166-174).  The propagation kernels never call back into Python:
they receive ``kind`` and the arrays through the plan (qc_set_static).
"""
from __future__ import annotations

import numpy

from . import engine as _eng
from .psi_basis import Psi


class Hamil2D:
    kind = None          # engine hamil code

    def __init__(self, v, akx2, np, U, W, delta, ntriv):
        self._v = v
        self._akx2 = akx2
        self._np = np
        self._U = U
        self._W = W
        self._delta = delta
        self._ntriv = ntriv

    # dense views used when a plan is built from this descriptor
    def potentials(self):
        return numpy.stack([numpy.asarray(vn[1], numpy.float64) for vn in self._v])

    def minima(self):
        return numpy.array([float(vn[0]) for vn in self._v])

    def kinetic(self):
        return numpy.ascontiguousarray(numpy.asarray(self._akx2).real)

    def __call__(self, orig: bool, psi: Psi, E, eL, E_full) -> Psi:
        from . import phys_base
        return phys_base.hamil_apply(self, orig, psi, E, eL, E_full)


class Hamil2DNonTrivial(Hamil2D):
    kind = _eng.HAMIL_NONTRIV


class Hamil2DTwoLevels(Hamil2D):
    kind = _eng.HAMIL_TWO_LEVELS


class Hamil2DBHX(Hamil2D):
    kind = _eng.HAMIL_BH_X


class Hamil2DBHZ(Hamil2D):
    kind = _eng.HAMIL_BH_Z


class Hamil2DQubit(Hamil2D):
    kind = _eng.HAMIL_QUBIT
