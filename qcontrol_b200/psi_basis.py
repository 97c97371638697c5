"""State containers with the reference's interface (psi_basis.py:7-53).

``Psi.f`` is a list of ``lvls`` complex128 arrays (one per level), ``PsiBasis.psis`` a list of ``n``
``Psi``; both deep-copy their arrays.  ``as_array`` / ``from_array`` convert to the dense
(nb, nlevs, np) layout the CUDA plans use.
"""
from __future__ import annotations

import copy
from typing import List, Optional, Sequence

import numpy


class Psi:
    __slots__ = ("_levels",)

    def __init__(self, f: Optional[Sequence[numpy.ndarray]] = None, lvls: int = 2):
        if f is None:
            self._levels = [numpy.zeros(1, dtype=numpy.complex128) for _ in range(lvls)]
        else:
            self._levels = [f[k] for k in range(lvls)]

    @property
    def f(self) -> List[numpy.ndarray]:
        return self._levels

    def lvls(self) -> int:
        return len(self._levels)

    def __deepcopy__(self, memo):
        return Psi(f=[a.copy() for a in self._levels], lvls=len(self._levels))

    def as_array(self) -> numpy.ndarray:
        return numpy.stack([numpy.asarray(a, numpy.complex128) for a in self._levels])

    @staticmethod
    def from_array(arr) -> "Psi":
        return Psi(f=[numpy.array(row, numpy.complex128) for row in arr], lvls=len(arr))


class PsiBasis:
    def __init__(self, n: int, lvls: int = 2):
        self.n = n
        self.lvls = lvls
        self._vectors = [Psi(lvls=lvls) for _ in range(n)]

    @property
    def psis(self) -> List[Psi]:
        return self._vectors

    def __len__(self):
        return len(self._vectors)

    def __deepcopy__(self, memo):
        out = PsiBasis(self.n, self.lvls)
        for k, p in enumerate(self._vectors):
            out._vectors[k] = copy.deepcopy(p, memo)
        return out

    def as_array(self) -> numpy.ndarray:
        return numpy.stack([p.as_array() for p in self._vectors])

    @staticmethod
    def from_array(arr) -> "PsiBasis":
        out = PsiBasis(len(arr), len(arr[0]))
        for k in range(len(arr)):
            out._vectors[k] = Psi.from_array(arr[k])
        return out
