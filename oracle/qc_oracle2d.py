"""CPU restatement of the two-dimensional extension (kernel C) -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference (irenemizus/qcontrol) has no two-dimensional operator ("hamil_2d 512x512" of
BASELINE config 5 names a grid the reference cannot run).  This module is the 2-D analogue
of the reference's 1-D path, built from the same pieces in the same order:

    diff      phys_base.diff_cpu    phys_base.py:18-36     ifft(fft(psi) * akx2)        -> along x and along y
    hamil     phys_base.hamil_cpu   phys_base.py:39-70     T psi + v psi                -> T_x psi + T_y psi + V psi
    residum   phys_base.residum_cpu phys_base.py:73-113    c1 H phi - c2 phi - xp phi   (unchanged)
    prop      phys_base.prop_cpu    phys_base.py:129-178   Newton polynomial + phase    (unchanged; coefficients from
                                                           oracle.qc_oracle.points = math_base.points)
    step      propagation.py:366-459  t_sc, renormalisation by sqrt(|<psi|psi>|)       (dx -> dx*dy)

What ties it to the reference nevertheless: for a separable potential V(x,y) = V1(x) + V2(y) and a product state
the exact propagator factorises, so a CONVERGED 2-D step must equal the outer product of two 1-D steps of the
pinned 1-D oracle (tests/test_oracle.py::test_grid2d_separable).
"""
from __future__ import annotations

import cmath

import numpy as np

from . import qc_oracle as qo


def apply_hamil_2d(psi, v, akx2, aky2):
    """H psi = ifft_x(akx2 fft_x psi) + ifft_y(aky2 fft_y psi) + v psi on psi[x][y]."""
    tx = np.fft.ifft(np.fft.fft(psi, axis=0) * akx2[:, None], axis=0)
    ty = np.fft.ifft(np.fft.fft(psi, axis=1) * aky2[None, :], axis=1)
    return tx + ty + v * psi


def newton_propagate_2d(psi, v, akx2, aky2, nch, t_sc, emin, emax):
    xp, dv = qo.points(nch, t_sc)
    c1 = 4.0 / (emax - emin)
    c2 = 2.0 * (emax + emin) / (emax - emin)
    phi = psi.copy()
    out = psi * dv[0]
    for j in range(nch - 1):
        phi = c1 * apply_hamil_2d(phi, v, akx2, aky2) - c2 * phi - xp[j] * phi
        out = out + dv[j + 1] * phi
    return out * cmath.exp(-1j * 2.0 * t_sc * (emax + emin) / (emax - emin))


def energy_window_2d(v, akx2, aky2, n):
    """propagation.py:371-375 with both kinetic maxima."""
    kmax = abs(akx2[int(n / 2 - 1)]) + abs(aky2[int(n / 2 - 1)])
    return float(v.max() + kmax), float(v.min())


def step_2d(psi, v, akx2, aky2, nch, dt, dxdy, emin, emax, renorm=True):
    t_sc = dt * (emax - emin) * qo.CM_TO_ERG / 4.0 / qo.RED_PLANCK_H
    out = newton_propagate_2d(psi, v, akx2, aky2, nch, t_sc, emin, emax)
    nrm = np.vdot(out, out) * dxdy
    if renorm and abs(nrm) > 0.0:
        out = out / np.sqrt(abs(nrm))
    return out, float(abs(nrm))
