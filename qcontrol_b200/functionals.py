"""The three control functionals of the reference and their 'a' coefficients -- the ONE definition used by
``task_manager`` (the reference-shaped static methods handed to ``FittingSolver``) and by the batched control
loops in ``control`` (task_manager.py:162-268 of the reference).

    F_sm = -|sum_v gc_v|^2      a_sm[v] = sum_{v', n} <chi_{v'n}(0)|psi_init_{v'n}> dx   (the same for every v)
    F_re = -sum_v Re gc_v       a_re[v] = 0.5
    F_ss = -sum_v |gc_v|^2      a_ss[v] = sum_n <chi_{vn}(0)|psi_init_{vn}> dx

``F_value`` / ``a_value`` keep the reference's loop and summation order (bit-identical results on one run);
``F_batch`` / ``a_batch`` are the same formulas over a whole batch of runs at once (numpy reductions, last-bit
differences in the summation order) for the throughput mode of big batches.
"""
from __future__ import annotations

import numpy

KINDS = ("sm", "re", "ss")


def F_value(kind, gc_vec, nb):
    """task_manager.py:162-207 on one run's goal-closeness vector gc_vec[nb]."""
    F = numpy.complex128(0)
    if kind == "sm":
        for a in range(nb):
            for b in range(nb):
                F -= gc_vec[a] * gc_vec[b].conjugate()
    elif kind == "re":
        for a in range(nb):
            F -= gc_vec[a].real
    elif kind == "ss":
        for a in range(nb):
            F -= gc_vec[a] * gc_vec[a].conjugate()
    else:
        raise KeyError(kind)
    return F


def a_value(kind, psi_init, chi_init, dx):
    """task_manager.py:209-268 on one run: psi_init, chi_init are arrays [nb][nlevs][np]."""
    nb, nlevs, _ = psi_init.shape
    a = numpy.zeros(nb, dtype=numpy.complex128)
    if kind == "re":
        a[:] = numpy.complex128(0.5)
    elif kind == "sm":
        for va in range(nb):
            for v in range(nb):
                for n in range(nlevs):
                    a[va] += numpy.vdot(chi_init[v, n], psi_init[v, n]) * dx       # math_base.cprod(psi, chi, dx, np)
    elif kind == "ss":
        for v in range(nb):
            for n in range(nlevs):
                a[v] += numpy.vdot(chi_init[v, n], psi_init[v, n]) * dx
    else:
        raise KeyError(kind)
    return a


def F_batch(kind, gc):
    """Re F of every run of a batch: gc[batch][nb] -> float64[batch]."""
    if kind == "sm":
        tau = gc.sum(axis=1)
        return -(tau * numpy.conj(tau)).real
    if kind == "re":
        return -gc.real.sum(axis=1)
    if kind == "ss":
        return -(gc * numpy.conj(gc)).real.sum(axis=1)
    raise KeyError(kind)


def a_batch(kind, psi_init, chi_init, dx):
    """'a' coefficients of every run: psi_init, chi_init [batch][nb][nlevs][np], dx[batch] -> complex128[batch][nb]."""
    per_v = (numpy.conj(chi_init) * psi_init).sum(axis=(2, 3)) * numpy.asarray(dx)[:, None]
    if kind == "sm":
        return numpy.repeat(per_v.sum(axis=1, keepdims=True), per_v.shape[1], axis=1)
    if kind == "re":
        return numpy.full_like(per_v, 0.5)
    if kind == "ss":
        return per_v
    raise KeyError(kind)


def kind_of(F_type):
    """Which functional a reference-style callable ``F_type(gc_vec, nb)`` is, found by probing it."""
    tag = getattr(F_type, "kind", None)
    if tag in KINDS:
        return tag
    probe = numpy.array([1.0 + 2.0j, 0.5 - 1.0j])
    val = F_type(probe, 2)
    for kind in KINDS:
        if abs(val - F_value(kind, probe, 2)) < 1e-12:
            return kind
    raise ValueError("unknown functional")
