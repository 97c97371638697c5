"""Host-side exact precompute for the propagation kernels.

Same public names and argument meaning as the reference's ``math_base.py``
(cprod/cprod2/cprod3 math_base.py:7-59, initak 62-91, reorder 119-150,
points 153-188).  The interpolation nodes and divided differences are INPUTS of
the CUDA kernels, so they are evaluated here in float64/complex128 with the
reference's operation order (the order decides the round-off-sized tail of the
coefficients) and uploaded; everything else in this file is O(np)
setup arithmetic.
"""
from __future__ import annotations

import cmath
import math
from functools import lru_cache

import numpy

__all__ = ["cprod", "cprod2", "cprod3", "initak", "reorder", "points", "newton_table"]


def cprod(cx1, cx2, dx, np):
    """<cx1|cx2> * dx with the reference's conjugation: dx * sum(conj(cx2) * cx1)."""
    if cx1.size != np or cx2.size != np:
        raise AssertionError("cprod: vectors must have np elements")
    return numpy.vdot(cx2, cx1) * dx


def cprod2(cx1, cx, dx, np):
    """<cx1| cx |cx1> * dx for a multiplicative operator cx."""
    if cx1.size != np or cx.size != np:
        raise AssertionError("cprod2: vectors must have np elements")
    return numpy.vdot(cx1, cx1 * cx) * dx


def cprod3(cx1, cx, cx2, dx, np):
    """dx * sum(conj(cx2) * cx1 * cx)."""
    if cx1.size != np or cx.size != np or cx2.size != np:
        raise AssertionError("cprod3: vectors must have np elements")
    return numpy.vdot(cx2, cx1 * cx) * dx


def initak(n, dx, iorder, ntriv):
    """(i k)^iorder on the FFT frequency grid, dk = 2 pi / ((n-1) dx); zeros when ntriv <= 0."""
    ak = numpy.zeros(n, numpy.complex128)
    if ntriv <= 0:
        return ak
    dk = 2.0 * math.pi / (n - 1) / dx
    mirror = pow(-1, iorder)
    for m in range(1, n // 2 + 1):
        val = pow(1j * dk * numpy.float64(m), iorder)
        ak[m] = val
        ak[n - m] = mirror * val
    return ak


@lru_cache(maxsize=None)
def _fold_order(nch):
    if nch <= 0 or nch & (nch - 1):
        raise AssertionError("nch must be a positive power of two")
    bits = nch.bit_length() - 1
    # repeatedly stacking the right half of a row under its left half and reading the first
    # column is the bit reversal of the row index
    return tuple(int(format(i, "0%db" % bits)[::-1], 2) if bits else 0 for i in range(nch))


def reorder(nch):
    """Visiting order of the Chebyshev nodes (list of ints), e.g. reorder(8) = [0,4,2,6,1,5,3,7]."""
    return list(_fold_order(nch))


def points(nch, t, func):
    """Nodes xp (fold order, on [-2, 2]) and Newton divided differences dv of func(z, t)."""
    order = _fold_order(nch)
    xp = numpy.empty(nch, numpy.float64)
    for i, j in enumerate(order):
        xp[i] = 2.0 * math.cos(numpy.float64(2 * j + 1) / numpy.float64(2 * nch) * math.pi)
    dv = numpy.zeros(nch, numpy.complex128)
    dv[0] = func(xp[0], t)
    if nch > 1:
        dv[1] = (func(xp[1], t) - dv[0]) / (xp[1] - xp[0])
    for i in range(2, nch):
        prod = 1.0
        tail = numpy.complex128(0)
        xi = xp[i]
        for j in range(i - 1):
            prod *= (xi - xp[j])
            tail += dv[j + 1] * prod
        prod *= (xi - xp[i - 1])
        dv[i] = (func(xi, t) - dv[0] - tail) / prod
    return xp, dv


_TABLE_CACHE = {}


def newton_table(nch, t_sc):
    """points(nch, t_sc, exp(-i z t)) memoised on (nch, t_sc): the reference rebuilds the table on
    every step (phys_base.py:155) although it only depends on these two numbers."""
    key = (int(nch), float(t_sc))
    hit = _TABLE_CACHE.get(key)
    if hit is None:
        if len(_TABLE_CACHE) > 4096:
            _TABLE_CACHE.clear()
        hit = points(nch, t_sc, lambda z, t: cmath.exp(-1j * z * t))
        _TABLE_CACHE[key] = hit
    return hit


def newton_tables_batched(nch, t_sc):
    """points() for a vector of scaled times at once (one row of dv per entry of t_sc).

    Same recurrence and operation order as points(), evaluated with numpy array arithmetic over the
    batch axis; used to set up large synthetic batches where a Python-scalar table per run would
    dominate the set-up time."""
    t = numpy.asarray(t_sc, numpy.float64).reshape(-1)
    order = _fold_order(nch)
    xp = numpy.array([2.0 * math.cos(numpy.float64(2 * j + 1) / numpy.float64(2 * nch) * math.pi) for j in order])
    f = numpy.exp(-1j * xp[None, :] * t[:, None])
    dv = numpy.zeros((t.size, nch), numpy.complex128)
    dv[:, 0] = f[:, 0]
    if nch > 1:
        dv[:, 1] = (f[:, 1] - dv[:, 0]) / (xp[1] - xp[0])
    for i in range(2, nch):
        prod = 1.0
        tail = numpy.zeros(t.size, numpy.complex128)
        for j in range(i - 1):
            prod *= (xp[i] - xp[j])
            tail += dv[:, j + 1] * prod
        prod *= (xp[i] - xp[i - 1])
        dv[:, i] = (f[:, i] - dv[:, 0] - tail) / prod
    return xp, dv


_PARTIAL_BOUND_CACHE = {}


def partial_product_bounds(nch):
    """B_j = max over z in [-2, 2] of |prod_{m<j} (z - xp_m)| for the nodes in the reference's visiting order: the
    norm of the j-th Newton basis operator prod_{m<j}(Hn - xp_m) for a Hermitian Hn with spectrum in [-2, 2]."""
    hit = _PARTIAL_BOUND_CACHE.get(nch)
    if hit is None:
        xp = numpy.array([2.0 * math.cos(numpy.float64(2 * j + 1) / numpy.float64(2 * nch) * math.pi)
                          for j in _fold_order(nch)])
        z = numpy.concatenate([numpy.linspace(-2.0, 2.0, 16001), xp])
        prod = numpy.ones_like(z)
        hit = numpy.empty(nch)
        for j in range(nch):
            hit[j] = numpy.abs(prod).max()
            prod = prod * (z - xp[j])
        hit *= 1.001                      # sampling margin
        _PARTIAL_BOUND_CACHE[nch] = hit
    return hit


def truncation_order(nch, dv, tol):
    """Smallest J >= 2 such that dropping the terms j >= J of the interpolation series changes P(Hn) psi by at most
    ``tol`` * |psi|:  sum_{j>=J} |dv_j| * B_j <= tol  (OPT-IN, see qc_set_poly_orders).  The reference sums all nch
    terms; behind the converged part its divided differences are round-off of size 1e-17, so for small
    scaled times most of the orders carry nothing: t_sc = 0.2 (single_harm) -> 10 of 64 at tol = 1e-11, t_sc = 3.7
    (the shipped Krotov input) -> 30 of 64; t_sc >= 14 needs them all."""
    dv = numpy.asarray(dv)
    B = partial_product_bounds(nch)
    tail = numpy.cumsum((numpy.abs(dv) * B)[..., ::-1], axis=-1)[..., ::-1]
    ok = tail <= tol
    J = numpy.where(ok.any(axis=-1), ok.argmax(axis=-1), nch)
    return numpy.maximum(J, 2).astype(numpy.int32)
