"""CPU oracle for the qcontrol propagation hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a from-scratch numpy restatement of the reference algorithm
(irenemizus/qcontrol, pure Python + numpy).  It exists so that the CUDA path
can be checked on a machine where ``/root/reference`` does not exist.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it; the product package
``qcontrol_b200`` never does (it fails loudly when its CUDA library is absent).

Parity pinning: every function below is checked in ``tests/test_oracle.py``
against (a) the known-answer vectors of the reference's own unit tests
(``tests/math_base_tests.py:32-51``), (b) fixtures produced by running the
unmodified reference modules in the build container
(``tests/golden/gen_golden.py`` -> ``tests/golden/*.npz``) and (c) rows of the
reference's shipped golden tables (``test_data/*.py``).

State layout: a wavefunction ("Psi", psi_basis.py:7-27) is a complex128 array
of shape (nlevs, np); a basis ("PsiBasis", psi_basis.py:30-53) is
(nb, nlevs, np).  Every function cites the reference file:line it restates.
"""
from __future__ import annotations

import cmath
import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np

# ----------------------------------------------------------------------------
# unit constants                                             phys_base.py:11-15
# ----------------------------------------------------------------------------
HART_TO_CM = np.float64(219474.6313708)
CM_TO_ERG = np.float64(1.98644568e-16)
DALT_TO_AU = np.float64(1822.888486)
RED_PLANCK_H = np.float64(1.054572e-27)
HZ_TO_CM = np.float64(3.33564095e-11)

FORWARD = 1
BACKWARD = -1


# ----------------------------------------------------------------------------
# math_base.py
# ----------------------------------------------------------------------------
def cprod(a, b, dx):
    """<a|b> = dx * sum(conj(b) * a)                      math_base.py:7-21"""
    return np.vdot(b, a) * dx


def cprod3(a, c, b, dx):
    """dx * sum(conj(b) * (a*c))                          math_base.py:42-59"""
    return np.vdot(b, np.multiply(a, c)) * dx


def initak(n, dx, order, ntriv):
    """k-space multiplier (i k)^order, zero for trivial systems.  math_base.py:62-91"""
    ak = np.zeros(n, np.complex128)
    if ntriv > 0:
        dk = 2.0 * math.pi / (n - 1) / dx
        for i in range(n // 2):
            ak[i + 1] = pow(1j * dk * np.float64(i + 1), order)
            ak[n - i - 1] = pow(-1, order) * ak[i + 1]
    return ak


def reorder(nch):
    """Node visiting order ("fold" permutation).  math_base.py:94-150.

    Folding a 1 x nch row log2(nch) times (right half goes under the left
    half) and reading column 0 top to bottom is the bit-reversal permutation;
    it is built here by that same folding, on a plain list of rows.
    """
    assert nch > 0 and nch & (nch - 1) == 0
    rows = [list(range(nch))]
    while len(rows[0]) > 1:
        half = len(rows[0]) // 2
        rows = [r[:half] for r in rows] + [r[half:] for r in rows]
    return [r[0] for r in rows]


def func(z, t):
    """exp(-i z t)                                        phys_base.py:116-126"""
    return cmath.exp(-1j * z * t)


def points(nch, t):
    """Chebyshev nodes on [-2,2] in fold order + Newton divided differences of
    exp(-i z t).  Same arithmetic order as math_base.py:153-188."""
    jj = reorder(nch)
    xp = np.zeros(nch, np.float64)
    for i in range(nch):
        phase = np.float64(2 * jj[i] + 1) / np.float64(2 * nch) * math.pi
        xp[i] = 2.0 * math.cos(phase)
    dv = np.zeros(nch, np.complex128)
    dv[0] = func(xp[0], t)
    if nch > 1:
        dv[1] = (func(xp[1], t) - dv[0]) / (xp[1] - xp[0])
    for i in range(2, nch):
        res = 1.0
        acc = np.complex128(0)
        for j in range(i - 1):
            res *= (xp[i] - xp[j])
            acc += dv[j + 1] * res
        res *= (xp[i] - xp[i - 1])
        dv[i] = (func(xp[i], t) - dv[0] - acc) / res
    return xp, dv


# ----------------------------------------------------------------------------
# laser envelopes and carriers                          task_manager.py:18-159
# ----------------------------------------------------------------------------
def env_zero(E0, t, t0, sigma):
    return np.float64(0.0)


def env_const(E0, t, t0, sigma):
    return E0


def env_gauss(E0, t, t0, sigma):
    return E0 * math.exp(-(t - t0) * (t - t0) / 2.0 / sigma / sigma)


def env_sqrsin(E0, t, t0, sigma):
    return E0 * math.sin(2.0 * math.pi * (t - t0) / sigma) * math.sin(2.0 * math.pi * (t - t0) / sigma)


def env_maxwell(E0, t, t0, sigma):
    return E0 * t * t * math.exp(-(t - t0) * (t - t0) / 2.0 / sigma / sigma) / sigma / sigma


ENVELOPES = {"zero": env_zero, "const": env_const, "gauss": env_gauss,
             "sqrsin": env_sqrsin, "maxwell": env_maxwell}


def hf_exp(nu, fm, t, pcos, w):
    return np.complex128(cmath.exp(1j * 2.0 * math.pi * nu * fm * t))


def hf_cos(nu, fm, t, pcos, w):
    return np.float64(math.cos(2.0 * math.pi * pcos * nu * fm * t))


def hf_sin(nu, fm, t, pcos, w):
    return np.float64(math.sin(2.0 * math.pi * pcos * nu * fm * t))


def hf_cos_set(nu, fm, t, pcos, w):
    out = w[0] * math.cos(2.0 * math.pi * nu * fm * t)
    i = -1
    for p in range(2, math.floor(pcos) + 1):
        i += 2
        out += w[i] * math.cos(2.0 * math.pi * nu * fm * t * p)
        out += w[i + 1] * math.cos(2.0 * math.pi * nu * fm * t / p)
    return np.float64(out)


def hf_sin_set(nu, fm, t, pcos, w):
    out = w[0] * math.sin(2.0 * math.pi * nu * fm * t)
    for p in range(1, math.floor(pcos) + 1):
        out += w[p] * math.sin(2.0 * math.pi * nu * fm * t * (2 * p + 1))
    return np.float64(out)


CARRIERS = {"exp": hf_exp, "cos": hf_cos, "sin": hf_sin,
            "cos_set": hf_cos_set, "sin_set": hf_sin_set}


# ----------------------------------------------------------------------------
# functionals                                           task_manager.py:162-268
# ----------------------------------------------------------------------------
def F_sm(gc, nb):
    F = np.complex128(0)
    for a in range(nb):
        for b in range(nb):
            F -= gc[a] * gc[b].conjugate()
    return F


def F_re(gc, nb):
    F = np.complex128(0)
    for a in range(nb):
        F -= gc[a].real
    return F


def F_ss(gc, nb):
    F = np.complex128(0)
    for a in range(nb):
        F -= gc[a] * gc[a].conjugate()
    return F


def a_sm(psi_init, chi_init, dx):
    nb, nlevs, _ = psi_init.shape
    a = np.zeros(nb, np.complex128)
    for va in range(nb):
        for v in range(nb):
            for n in range(nlevs):
                a[va] += cprod(psi_init[v, n], chi_init[v, n], dx)
    return a


def a_re(psi_init, chi_init, dx):
    return np.full(psi_init.shape[0], np.complex128(0.5))


def a_ss(psi_init, chi_init, dx):
    nb, nlevs, _ = psi_init.shape
    a = np.zeros(nb, np.complex128)
    for v in range(nb):
        for n in range(nlevs):
            a[v] += cprod(psi_init[v, n], chi_init[v, n], dx)
    return a


FUNCTIONALS = {"sm": (F_sm, a_sm), "re": (F_re, a_re), "ss": (F_ss, a_ss)}


# ----------------------------------------------------------------------------
# the problem description (what task_manager.TaskManager produces)
# ----------------------------------------------------------------------------
@dataclass
class Problem:
    """Everything FittingSolver/PropagationSolver receive.  Built by
    ``make_problem`` (restating task_manager.py:350-517, 541-948) or by hand."""
    task_type: str = "trans_wo_control"
    hamil: str = "ntriv"            # ntriv | two_levels | bh_x | bh_z | qubit
    ntriv: int = 1
    np_: int = 1024
    nlevs: int = 2
    nb: int = 1
    L: float = 5.0
    T: float = 600e-15
    dx: float = 0.0
    x: np.ndarray = None
    v: np.ndarray = None            # (nlevs, np) float64
    vmin: np.ndarray = None         # (nlevs,)   the "v[n][0]" entries
    akx2: np.ndarray = None         # (np,) complex128
    psi0: np.ndarray = None         # (nb, nlevs, np)
    psif: np.ndarray = None
    U: float = 0.0
    W: float = 0.0
    delta: float = 1.0
    m: float = 0.5
    nch: int = 64
    nt: int = 420000
    E0: float = 71.54
    t0: float = 300e-15
    sigma: float = 50e-15
    nu_L: float = 0.29297e15        # conf.fitter.propagation.nu_L  (eL uses this)
    nu: float = 0.29297e15          # task-manager carrier frequency (nu_L_auto aware)
    envelope: str = "zero"
    carrier: str = "exp"
    pcos: float = 1.0
    w_list: Sequence[float] = ()
    hf_hide: bool = True
    h_lambda: float = 0.0066
    h_lambda_mode: str = "const"
    epsilon: float = 1e-15
    iter_max: int = -1
    iter_mid_1: int = 0
    iter_mid_2: int = 1
    q: float = 0.0
    impulses_number: int = 2
    delay: float = 600e-15
    k_E: float = 1e29               # local control: spring constant of the field / multiplier ODE [1/s^2]
    lamb: float = 4e14              # local control: decay rate [1/s]
    pow_: float = 0.8               # local control: exponent of the decay term
    F_type: str = "sm"
    F_goal: float = -1.0
    init_dir: int = FORWARD
    t_step: float = 0.0
    t_list: List[float] = field(default_factory=list)

    def laser_field(self, E0, t, t0, sigma):
        return ENVELOPES[self.envelope](E0, t, t0, sigma)

    def laser_field_hf(self, freq_mult, t, pcos, w_list):
        return CARRIERS[self.carrier](self.nu, freq_mult, t, pcos, w_list)


# ----------------------------------------------------------------------------
# Hamiltonian application                    phys_base.py:18-70, hamil_2d.py:33-187
# ----------------------------------------------------------------------------
def diff(psi_n, akx2):
    """ifft(fft(psi) * akx2)                              phys_base.py:18-36"""
    return np.fft.ifft(np.multiply(np.fft.fft(psi_n), akx2))


def hamil_1d(psi_n, v_n, akx2, ntriv):
    """phys_base.py:39-70"""
    if ntriv == 1:
        phi = diff(psi_n, akx2)
        np.add(phi, np.multiply(v_n, psi_n), out=phi)
        return phi
    return np.multiply(v_n, psi_n)


def _jcoef(l, vi):
    return math.sqrt(l * (l + 1) - (l - vi + 1) * (l - vi))


def level_matrix(pb: Problem, E, E_full):
    """The nlevs x nlevs matrix built per call by BH_X / BH_Z / qubit.
    hamil_2d.py:100-118, 130-148, 160-178."""
    nl = pb.nlevs
    l = (nl - 1) / 2.0
    H = np.zeros((nl, nl), np.complex128)
    U, W, d = pb.U, pb.W, pb.delta
    if pb.hamil == "bh_x":
        H[0, 0] = 2.0 * l * U + 2.0 * l * l * W
        for vi in range(1, nl):
            H[vi, vi] = 2.0 * (l - vi) * U + 2.0 * (l - vi) * (l - vi) * W
            H[vi - 1, vi] = -d * E * _jcoef(l, vi)
            H[vi, vi - 1] = -d * E * _jcoef(l, vi)
    elif pb.hamil == "bh_z":
        H[0, 0] = 2.0 * l ** 2 * U + 2.0 * l * E
        for vi in range(1, nl):
            H[vi, vi] = 2.0 * (l - vi) ** 2 * U + 2.0 * (l - vi) * E
            H[vi - 1, vi] = -d * _jcoef(l, vi)
            H[vi, vi - 1] = -d * _jcoef(l, vi)
    elif pb.hamil == "qubit":
        H[0, 0] = -2.0 * l * U + 2.0 * l * l * W
        for vi in range(1, nl):
            H[vi, vi] = -2.0 * (l - vi) * U + 2.0 * (l - vi) * (l - vi) * W
            H[vi - 1, vi] = -d * E_full * _jcoef(l, vi)
            H[vi, vi - 1] = -d * np.conjugate(E_full) * _jcoef(l, vi)
    else:
        raise ValueError(pb.hamil)
    return H


def apply_hamil(pb: Problem, psi, E, eL, E_full, orig):
    """phi = H psi for one wavefunction (nlevs, np).  hamil_2d.py:33-187."""
    if pb.hamil == "ntriv":                               # hamil_2d.py:34-73
        if orig:
            s_l = 0.0
            s_u = 0.0
            c_l = psi[0] * np.conjugate(E_full)
            c_u = psi[1] * E_full
        else:
            s_l = psi[0] * eL
            s_u = psi[1] * eL
            c_l = psi[0] * E
            c_u = psi[1] * E
        d_l = hamil_1d(psi[0], pb.v[0], pb.akx2, pb.ntriv)
        d_u = hamil_1d(psi[1], pb.v[1], pb.akx2, pb.ntriv)
        np.add(d_l, s_l, out=d_l)
        np.subtract(d_u, s_u, out=d_u)
        return np.stack([np.subtract(d_l, c_u), np.subtract(d_u, c_l)])
    if pb.hamil == "two_levels":                          # hamil_2d.py:76-97
        c_l = psi[0] * np.conjugate(E_full)
        c_u = psi[1] * E_full
        d_l = hamil_1d(psi[0], pb.v[0], pb.akx2, pb.ntriv)
        d_u = hamil_1d(psi[1], pb.v[1], pb.akx2, pb.ntriv)
        return np.stack([np.subtract(d_l, c_u), np.subtract(d_u, c_l)])
    H = level_matrix(pb, E, E_full)                       # hamil_2d.py:120-125
    out = np.zeros_like(psi)
    for g in range(pb.nlevs):
        acc = np.zeros(psi.shape[1], np.complex128)
        for i in range(pb.nlevs):
            acc = np.add(acc, H[g, i] * psi[i])
        out[g] = acc
    return out


def residum(pb, phi, xp_j, emin, emax, E, eL, E_full, orig):
    """(Hn - xp_j) phi with Hn scaled to [-2,2].          phys_base.py:73-113"""
    h = apply_hamil(pb, phi, E, eL, E_full, orig)
    c1 = 4.0 / (emax - emin)
    c2 = 2.0 * (emax + emin) / (emax - emin)
    out = np.empty_like(phi)
    for n in range(phi.shape[0]):
        hn = h[n]
        hn *= c1
        np.subtract(hn, phi[n] * c2, out=hn)
        pn = phi[n] * (-xp_j)
        np.add(pn, hn, out=pn)
        out[n] = pn
    return out


def newton_propagate(pb, psi, t_sc, emin, emax, E, eL, E_full, orig=False, coeffs=None):
    """One polynomial step psi <- P(Hn) psi.              phys_base.py:129-178

    ``coeffs`` lets callers reuse (xp, dv); the reference recomputes them on
    every call (phys_base.py:155) with identical results.
    """
    xp, dv = coeffs if coeffs is not None else points(pb.nch, t_sc)
    # Every array operation below acts on ONE level (a contiguous 1-D row) and is
    # in-place exactly where the reference is in-place (phys_base.py:162,176): numpy's
    # in-place and out-of-place complex multiply loops round differently (FMA vs
    # no FMA), and a bit-identical oracle has to follow that.
    phi = psi.copy()
    out = psi.copy()
    for n in range(out.shape[0]):
        out[n] *= dv[0]
    for j in range(pb.nch - 1):
        phi = residum(pb, phi, xp[j], emin, emax, E, eL, E_full, orig)
        for n in range(out.shape[0]):
            np.add(out[n], phi[n] * dv[j + 1], out=out[n])
    coef = cmath.exp(-1j * 2.0 * t_sc * (emax + emin) / (emax - emin))
    for n in range(out.shape[0]):
        out[n] *= coef
    return out


# ----------------------------------------------------------------------------
# instrumentation                      propagation.py:157-182, phys_base.py:196-288
# ----------------------------------------------------------------------------
def level_norms(psi, dx):
    return np.array([cprod(p, p, dx) for p in psi])


def level_energies(pb, psi, E, eL, E_full, orig=True):
    phi = apply_hamil(pb, psi, E, eL, E_full, orig)
    return np.array([cprod(psi[n], phi[n], pb.dx) for n in range(psi.shape[0])])


def level_overlaps(goal, psi, dx):
    return np.array([cprod(goal[n], psi[n], dx) for n in range(psi.shape[0])])


def moments(pb, psi):
    """<x>, <x^2>, <p>, <p^2> per level.                  phys_base.py:196-256"""
    nl = psi.shape[0]
    mx = np.zeros(nl, np.complex128)
    mx2 = np.zeros(nl, np.complex128)
    mp = np.zeros(nl, np.complex128)
    mp2 = np.zeros(nl, np.complex128)
    x = pb.x
    x2 = np.multiply(x, x)
    for n in range(nl):
        mx[n] = np.vdot(psi[n], np.multiply(psi[n], x)) * pb.dx
        mx2[n] = np.vdot(psi[n], np.multiply(psi[n], x2)) * pb.dx
    if pb.ntriv == 1:
        for n in range(nl):
            mp2[n] = cprod(psi[n], diff(psi[n], pb.akx2) * (2.0 * pb.m), pb.dx)
        akx = initak(pb.np_, pb.dx, 1, pb.ntriv)
        akx *= HART_TO_CM / (-1j) / DALT_TO_AU
        for n in range(nl):
            mp[n] = cprod(psi[n], diff(psi[n], akx), pb.dx)
    return mx, mx2, mp, mp2


def sigma_moments(pb, psi):
    """phys_base.py:259-288"""
    if pb.ntriv != 1 and psi.shape[0] == 2:
        c = cprod(psi[0], psi[1], pb.dx)
        sz = cprod(psi[0], psi[0], pb.dx) - cprod(psi[1], psi[1], pb.dx)
        return 2.0 * c.real, 2.0 * c.imag, sz
    return np.float64(0.0), np.float64(0.0), np.float64(0.0)


# ----------------------------------------------------------------------------
# time stepper                                        propagation.py:323-484
# ----------------------------------------------------------------------------
@dataclass
class StepRecord:
    l: int
    t: float
    E: complex
    freq_mult: float
    psi: np.ndarray
    cnorm: np.ndarray
    cener: np.ndarray
    overlp0: np.ndarray
    overlpf: np.ndarray
    emax: float
    emin: float
    t_sc: float
    psigc_psie: complex = 0.0
    psigc_dv_psie: complex = 0.0


class Stepper:
    """One basis vector's PropagationSolver (propagation.py:17-484), without
    reporters.  ``field_cb(stepper) -> E`` and ``freq_cb(stepper) -> freq_mult``
    are called at the same points as the reference's callbacks."""

    def __init__(self, pb: Problem, field_cb, freq_cb=None, instrument=True):
        self.pb = pb
        self.field_cb = field_cb
        self.freq_cb = freq_cb or (lambda st: np.float64(1.0))
        self.instrument = instrument
        self._coeff_cache = {}

    def energy_extremes(self):
        """propagation.py:371-375"""
        pb = self.pb
        extr = []
        for n in range(pb.nlevs):
            extr.append(pb.v[n][0] + abs(pb.akx2[int(pb.np_ / 2 - 1)]))
            extr.append(pb.vmin[n])
        return extr

    def start(self, psi0, psif, direction):
        """propagation.py:323-364 (no static report)."""
        pb = self.pb
        self.psi0 = psi0
        self.psif = psif
        self.dir = direction
        self.dt = direction * pb.t_step
        self.psi = psi0.copy()
        self.psi_omega = psi0.copy()
        self.l = 0
        self.t = np.float64(0.0) if direction == FORWARD else pb.T
        self.E = np.float64(0.0)
        self.freq_mult = np.float64(1.0)
        self.E = self.field_cb(self)
        self.l = 1
        self.rec = None

    def _coeffs(self, t_sc):
        key = float(t_sc)
        if key not in self._coeff_cache:
            if len(self._coeff_cache) > 8:
                self._coeff_cache.clear()
            self._coeff_cache[key] = points(self.pb.nch, t_sc)
        return self._coeff_cache[key]

    def step(self, t_start):
        """propagation.py:366-484"""
        pb = self.pb
        nl = pb.nlevs
        extr = self.energy_extremes()
        self.t = self.dt * self.l + t_start
        eL = pb.nu_L * self.freq_mult * HZ_TO_CM / 2.0
        exp_L = np.float64(1.0)
        if pb.hf_hide:
            self.freq_mult = self.freq_cb(self)
            exp_L = np.complex128(cmath.sqrt(pb.laser_field_hf(self.freq_mult, self.t, pb.pcos, pb.w_list)))
            self.psi_omega = np.stack([self.psi[0] / exp_L, self.psi[1] * exp_L])
            ex = [extr[0] + pb.E0 + eL, extr[1] - pb.E0 + eL,
                  extr[2] + pb.E0 - eL, extr[3] - pb.E0 - eL]
            emax, emin = max(ex), min(ex)
        else:
            emax, emin = max(extr), min(extr)
        t_sc = self.dt * (emax - emin) * CM_TO_ERG / 4.0 / RED_PLANCK_H
        self.E = self.field_cb(self)
        psigc_psie = np.float64(0.0)
        psigc_dv_psie = np.float64(0.0)
        coeffs = self._coeffs(t_sc)
        if pb.hf_hide:
            E_full = self.E * exp_L * exp_L
            po = newton_propagate(pb, self.psi_omega, t_sc, emin, emax, self.E, eL, E_full, False, coeffs)
            cnorm = [cprod(po[n], po[n], pb.dx) for n in range(nl)]
            s = 0.0
            for c in cnorm:
                s += c
            if abs(s) > 0.0:
                for n in range(nl):
                    po[n] /= math.sqrt(abs(s))
            psigc_psie = cprod(po[1], po[0], pb.dx)
            psigc_dv_psie = cprod3(po[1], pb.v[0] - pb.v[1], po[0], pb.dx)
            self.psi_omega = po
            self.psi = np.stack([po[0] * exp_L, po[1] / exp_L])
        else:
            E_full = self.E
            p = newton_propagate(pb, self.psi, t_sc, emin, emax, self.E, eL, E_full, False, coeffs)
            cnorm = [cprod(p[n], p[n], pb.dx) for n in range(nl)]
            s = 0.0
            for c in cnorm:
                s += c
            if abs(s) > 0.0:
                for n in range(nl):
                    p[n] /= math.sqrt(abs(s))
            self.psi = p                 # psi_omega stays as created in start()
        if self.instrument:
            cener = level_energies(pb, self.psi, self.E, eL, E_full, True)
            o0 = level_overlaps(self.psi0, self.psi, pb.dx)
            of = level_overlaps(self.psif, self.psi, pb.dx)
        else:
            cener = o0 = of = None
        self.rec = StepRecord(self.l, self.t, self.E, self.freq_mult, self.psi, np.array(cnorm), cener,
                              o0, of, emax, emin, t_sc, psigc_psie, psigc_dv_psie)
        self.l += 1
        return self.l <= pb.nt


# ----------------------------------------------------------------------------
# control loop                                              fitter.py:13-836
# ----------------------------------------------------------------------------
@dataclass
class IterRow:
    it: int
    goal_close: float
    Fsm: float
    E_int: float
    J: float
    E_tlist: np.ndarray


class Fitter:
    """Restatement of FittingSolver for every task type (single_pot,
    filtering, trans_wo_control, intuitive_control, local_control_population,
    local_control_projection, optimal_control_krotov,
    optimal_control_unit_transform).
    ``on_step(prop_id, vect, stepper)`` is called after every solver step
    (where the reference calls reporter.print_time_point_prop)."""

    def __init__(self, pb: Problem, on_step: Optional[Callable] = None, instrument=False):
        self.pb = pb
        self.on_step = on_step
        self.instrument = instrument or (on_step is not None)
        self.rows: List[IterRow] = []
        self.F_fn, self.a_fn = FUNCTIONALS[pb.F_type]

    # -- field callbacks ---------------------------------- fitter.py:650-811
    def _field_forward(self, st: Stepper):
        pb = self.pb
        if pb.hf_hide:
            patched = pb.laser_field(pb.E0, st.t, pb.t0, pb.sigma)
        else:
            patched = pb.laser_field(pb.E0, st.t, pb.t0, pb.sigma) * \
                pb.laser_field_hf(st.freq_mult, st.t, pb.pcos, pb.w_list)
        tt = pb.task_type
        self.E_patched = patched
        if tt == "intuitive_control":
            for k in range(1, pb.impulses_number):
                patched += pb.laser_field(pb.E0, st.t, pb.t0 + k * pb.delay, pb.sigma)
            self.E_patched = patched
            return patched
        if tt == "local_control_population":              # fitter.py:673-692
            if st.E == 0.0:
                st.E = self.E_patched
            if self.E_patched <= 0:
                raise RuntimeError("E_patched has to be positive")
            if st.E <= 0:
                raise RuntimeError("E has to be positive")
            first = st.E - self.E_patched
            second = self.E_vel * math.pow(self.E_patched / st.E, pb.pow_)
            E_acc = -pb.k_E * first - pb.lamb * second
            self.E_vel += E_acc * abs(st.dt)
            return st.E + self.E_vel * abs(st.dt)
        if tt == "optimal_control_krotov":
            if self.iter_step == 0:
                return patched
            ov = cprod(self.chi_tlist[pb.nt - st.l][0][1], st.psi_omega[0], pb.dx)
            return -2.0 * ov.imag / pb.h_lambda
        if tt == "optimal_control_unit_transform":
            if st.l == 0:
                chi_init = self.chi_tlist[-1]
                h0 = pb.h_lambda * RED_PLANCK_H / CM_TO_ERG if pb.ntriv == -1 else pb.h_lambda
                if pb.h_lambda_mode == "dynamical":
                    self.h_lambda = h0 * pb.nb / math.sqrt(self.goal_close_abs) if self.goal_close_abs else h0
                else:
                    self.h_lambda = h0
                self.a0 = self.a_fn(pb.psi0, chi_init, pb.dx)
                return patched
            chi = self.chi_tlist[-st.l]
            psi = self.psi_cur
            acc = np.float64(0.0)
            for v in range(pb.nb):
                for n in range(pb.nlevs):
                    acc += self.a0[v] * cprod(chi[v][n], psi[v][n], pb.dx)
            th = st.t - abs(st.dt) / 2.0
            hf = pb.laser_field_hf(st.freq_mult, th, pb.pcos, pb.w_list)
            s = pb.laser_field(pb.E0, th, pb.t0, pb.sigma) / pb.E0
            dE = -s * acc.imag / self.h_lambda
            if self.iter_step == 0:
                return s * pb.E0 * hf + dE
            return self.E_tlist[st.l] + dE
        return patched

    def _field_backward(self, st: Stepper):
        pb = self.pb
        if pb.task_type == "optimal_control_krotov":
            ov = cprod(st.psi_omega[1], self.psi_tlist[pb.nt - st.l][0][0], pb.dx)
            return -2.0 * ov.imag / pb.h_lambda
        if pb.task_type == "optimal_control_unit_transform":
            if self.iter_step == 0:
                return pb.laser_field(pb.E0, st.t, pb.t0, pb.sigma) * \
                    pb.laser_field_hf(st.freq_mult, st.t, pb.pcos, pb.w_list)
            return self.E_tlist[-1] if st.l == 0 else self.E_tlist[-st.l]
        return pb.laser_field(pb.E0, st.t, pb.t0, pb.sigma)

    def _freq_mult(self, st: Stepper):
        """fitter.py:814-836"""
        pb = self.pb
        if pb.task_type == "local_control_projection":
            if self.fm_patched < 0:
                raise RuntimeError("freq_mult_patched has to be positive or zero")
            if st.freq_mult < 0:
                raise RuntimeError("Frequency multiplier has to be positive or zero")
            first = st.freq_mult - self.fm_patched
            second = self.fm_vel * math.pow(self.fm_patched / st.freq_mult, pb.pow_)
            acc = -pb.k_E * first - pb.lamb * second
            self.fm_vel += acc * abs(st.dt)
            return np.float64(st.freq_mult + self.fm_vel * abs(st.dt))
        return np.float64(self.fm_patched)

    def _feedback(self, st: Stepper):
        """do_the_thing, fitter.py:578-623 (the state it leaves behind; no printing)."""
        pb = self.pb
        r = st.rec
        if pb.task_type == "local_control_population":
            coef = 2.0 * CM_TO_ERG / RED_PLANCK_H
            dAdt = st.E * r.psigc_psie.imag * coef
            if dAdt >= 0.0:
                self.dAdt_happy = dAdt
            elif abs(r.psigc_psie.imag) > pb.epsilon:
                self.E_patched = -self.dAdt_happy / (r.psigc_psie.imag * coef)
            else:
                self.E_patched = self.dAdt_happy / (pb.epsilon * coef)
            self.dAdt = dAdt
        elif pb.task_type == "local_control_projection":
            coef2 = -4.0 * CM_TO_ERG / RED_PLANCK_H
            Sge2 = r.psigc_psie * r.psigc_psie
            Sdvge = r.psigc_psie * r.psigc_dv_psie
            freq_cm = HZ_TO_CM * pb.nu_L
            body = Sdvge + freq_cm * st.freq_mult * Sge2
            dAdt = body.imag * coef2
            if dAdt >= 0.0:
                self.dAdt_happy = dAdt
            else:
                if Sge2.imag > pb.epsilon:
                    self.fm_patched = (self.dAdt_happy - coef2 * Sdvge.imag) / (Sge2.imag * freq_cm * coef2)
                elif Sge2.imag > 0.0:
                    self.fm_patched = (self.dAdt_happy - coef2 * Sdvge.imag) / (pb.epsilon * freq_cm * coef2)
                elif Sge2.imag < -pb.epsilon:
                    self.fm_patched = -Sdvge.imag / (Sge2.imag * freq_cm)
                else:
                    self.fm_patched = Sdvge.imag / (pb.epsilon * freq_cm)
                if self.fm_patched < 0.0:
                    self.fm_patched = np.float64(0.0)
            self.dAdt = dAdt

    # -- one sweep ---------------------------------------- fitter.py:213-422
    def _sweep(self, direction, chiT):
        pb = self.pb
        nb = pb.nb
        self.dir = direction
        self.chi_cur = None
        self.psi_cur = None
        if direction == FORWARD:
            tag = "f"
            cb = self._field_forward
            self.psi_tlist = [pb.psi0.copy()]
            self.psi_cur = self.psi_tlist[0]
            init, fin = pb.psi0, pb.psif
            t_init = np.float64(0.0)
        else:
            tag = "b"
            cb = self._field_backward
            if pb.task_type == "optimal_control_unit_transform":
                chiT = pb.psif.copy()
                self.chi_tlist = [chiT]
            self.chi_cur = self.chi_tlist[0]
            init, fin = chiT, pb.psi0
            t_init = pb.T
        prop_id = "iter_%d%s" % (self.iter_step, tag)
        solvers = [Stepper(pb, cb, freq_cb=self._freq_mult, instrument=self.instrument) for _ in range(nb)]
        for v in range(nb):
            solvers[v].start(init[v], fin[v], direction)
        E_new = [solvers[0].E] if direction == FORWARD else []
        more = True
        while more:
            new = np.empty((nb, pb.nlevs, pb.np_), np.complex128)
            alive = 0
            for v in range(nb):
                if solvers[v].step(t_init):
                    alive += 1
                new[v] = solvers[v].psi_omega if pb.hf_hide else solvers[v].psi
                if self.on_step:
                    self.on_step(prop_id, v, solvers[v])
                self._feedback(solvers[v])
            if direction == BACKWARD:
                self.chi_cur = new
                self.chi_tlist.append(new)
            else:
                E_new.append(solvers[0].E)
                self.psi_cur = new
                if pb.task_type == "optimal_control_krotov":
                    self.psi_tlist.append(new)
            more = alive == nb
        if direction == FORWARD:
            self.E_tlist = E_new
            self.E_int = np.float64(0.0)
            for k in range(len(E_new) - 1):
                Ec = E_new[k + 1]
                self.E_int += Ec * np.conjugate(Ec) * (pb.t_step * 1e15)
        gc = np.zeros(nb, np.complex128)
        chiT_omega = np.zeros((nb, pb.nlevs, pb.np_), np.complex128)
        chiT_out = chiT
        for v in range(nb):
            sv = solvers[v]
            if direction == FORWARD:
                for n in range(pb.nlevs):
                    gc[v] += cprod(sv.psif[n], sv.psi[n], pb.dx)
                if pb.task_type == "optimal_control_krotov":
                    part = np.zeros((pb.nlevs, pb.np_), np.complex128)
                    part[1] = gc[v] * sv.psif[1]
                    cn = cprod(part[1], part[1], pb.dx)
                    if abs(cn) > 0.0:
                        part[1] /= math.sqrt(abs(cn))
                    po = part.copy()
                    po[1] *= cmath.sqrt(pb.laser_field_hf(sv.freq_mult, pb.T, pb.pcos, pb.w_list))
                    chiT_omega[v] = po
                    if chiT_out is None:
                        chiT_out = np.zeros_like(chiT_omega)
                    chiT_out[v] = part
            self.goal_close_scal += gc[v]
        self.goal_close_vec = gc
        if pb.task_type == "optimal_control_krotov" and direction == FORWARD:
            self.chi_tlist = [chiT_omega]
        if direction == FORWARD:
            self.goal_close_abs = abs(self.goal_close_scal)
        self.Fsm = self.F_fn(gc, nb)
        if pb.task_type == "optimal_control_krotov":
            hl = pb.h_lambda
        elif pb.task_type == "optimal_control_unit_transform":
            hl = self.h_lambda
        else:
            hl = np.float64(0.0)
        self.J = self.Fsm.real - hl * hl * np.real(self.E_int)
        self.solvers = solvers
        return chiT_out

    def _report(self, it):
        E_list = np.array([abs(E) if np.imag(E) else np.real(E) for E in self.E_tlist], np.float64)
        self.rows.append(IterRow(it, float(self.goal_close_abs), float(np.real(self.Fsm)),
                                 float(np.real(self.E_int)), float(self.J), E_list))

    def _iteration_optimal(self):
        """fitter.py:441-489"""
        pb = self.pb
        chiT = np.zeros((pb.nb, pb.nlevs, pb.np_), np.complex128)
        self.goal_close_scal = np.float64(0.0)
        direction = pb.init_dir
        for _ in range(2):
            chiT = self._sweep(direction, chiT)
            if pb.task_type != "optimal_control_unit_transform":
                if abs(self.Fsm.real - pb.F_goal) <= pb.epsilon and self.iter_step == 0:
                    break
            if direction == FORWARD:
                self._report(self.iter_step)
            direction = -direction

    def run(self):
        """fitter.py:491-571"""
        pb = self.pb
        self.iter_step = 0
        self.h_lambda = np.float64(0.0)
        self.goal_close_abs = np.float64(0.0)
        self.goal_close_scal = np.float64(0.0)
        self.E_tlist = []
        self.E_int = np.float64(0.0)
        self.chi_tlist = []
        self.psi_tlist = []
        self.E_vel = np.float64(0.0)                      # FitterDynamicState, fitter.py:22-31
        self.fm_vel = np.float64(0.0)
        self.E_patched = np.float64(0.0)
        self.fm_patched = np.float64(1.0)
        self.dAdt_happy = np.float64(0.0)
        optimal = pb.task_type in ("optimal_control_krotov", "optimal_control_unit_transform")
        if pb.task_type == "optimal_control_unit_transform":
            gc0 = np.zeros(pb.nb, np.complex128)
            sc = np.float64(0.0)
            for v in range(pb.nb):
                for n in range(pb.nlevs):
                    gc0[v] += cprod(pb.psif[v][n], pb.psi0[v][n], pb.dx)
                sc += gc0[v]
            F0 = self.F_fn(gc0, pb.nb)
            E0l = []
            for t in pb.t_list:
                hf = pb.laser_field_hf(np.float64(1.0), t, pb.pcos, pb.w_list)
                E0l.append((pb.laser_field(pb.E0, t, pb.t0, pb.sigma) * hf).real)
            Ei = np.float64(0.0)
            for k in range(len(E0l) - 1):
                Ei += E0l[k + 1] * np.conjugate(E0l[k + 1]) * (pb.t_step * 1e15)
            h0 = pb.h_lambda * RED_PLANCK_H / CM_TO_ERG if pb.ntriv == -1 else pb.h_lambda
            self.rows.append(IterRow(-1, float(abs(sc)), float(F0.real), float(Ei),
                                     float(F0.real - h0 * h0 * Ei), np.array(E0l, np.float64)))
        if optimal:
            self._iteration_optimal()
            while abs(self.Fsm.real - pb.F_goal) > pb.epsilon and \
                    (self.iter_step < pb.iter_max or pb.iter_max == -1):
                self.iter_step += 1
                self._iteration_optimal()
                if pb.q:
                    if self.iter_step == pb.iter_mid_1:
                        self.F_mid_1 = self.Fsm.real
                    elif self.iter_step == pb.iter_mid_2:
                        self.F_mid_2 = self.Fsm.real
                        lim = -pb.nb * pb.nb * pb.q
                        if self.F_mid_1 > lim or self.F_mid_2 > lim or self.F_mid_1 < self.F_mid_2:
                            raise ValueError("The calculation is probably going to diverge.  Stopping the run...")
        else:
            self._sweep(FORWARD, pb.psif)
            self._report(self.iter_step)
        return self.rows


# ----------------------------------------------------------------------------
# task setup                   task_manager.py:279-337, 350-517, 541-948; grid_setup.py
# ----------------------------------------------------------------------------
def grid(L, n):
    """grid_setup.py:8-34"""
    if n > 1:
        dx = L / (n - 1)
        shift = np.float64(n - 1) * dx / 2.0
        return dx, np.array([np.float64(i) * dx - shift for i in range(n)])
    return np.float64(1.0), np.array([0.0])


def wf_harmonic(x, x0, p0, a):
    """task_manager.py:289-308"""
    return np.array([cmath.exp(-(xi - x0) * (xi - x0) / 2.0 / a / a + 1j * p0 * xi) /
                     pow(math.pi, 0.25) / math.sqrt(a) for xi in x]).astype(np.complex128)


def wf_morse(x, x0, m, De, a):
    """task_manager.py:311-337"""
    omega_0 = a * math.sqrt(2.0 * De / HART_TO_CM / m / DALT_TO_AU) * HART_TO_CM
    xe = omega_0 / 4.0 / De
    y = [math.exp(-a * (xi - x0)) / xe for xi in x]
    arg = 1.0 / xe - 1.0
    return np.array([math.sqrt(a / math.gamma(arg)) * math.exp(-yi / 2.0) * pow(yi, np.float64(arg / 2.0))
                     for yi in y]).astype(np.complex128)


def wf_one(x, L):
    """task_manager.py:285-286"""
    return np.array([1.0 / math.sqrt(L)] * len(x)).astype(np.complex128)


_DEFAULTS = dict(
    task_type="trans_wo_control", pot_type="morse", wf_type="morse", hamil_type="ntriv",
    lf_aug_type="z", init_guess="zero", init_guess_hf="exp", nb=1, nlevs=2, T=600e-15, L=5.0,
    np=1024, De=20000.0, De_e=10000.0, Du=20000.0, x0p=-0.17, x0=0.0, p0=0.0, a=1.0, a_e=1.0,
    U=0.0, W=0.0, delta=1.0, t0_auto=False, nt_auto=False, sigma_auto=False, nu_L_auto=False)
_FIT_DEFAULTS = dict(
    F_type="sm", w_list=[], k_E=1e29, lamb=4e14, pow=0.8, epsilon=1e-15, impulses_number=2,
    delay=600e-15, mod_log=500, iter_max=-1, iter_mid_1=0, iter_mid_2=1, q=0.0, h_lambda=0.0066,
    h_lambda_mode="const", pcos=1.0, hf_hide=True, Em=1.5)
_PROP_DEFAULTS = dict(m=0.5, nch=64, nt=420000, E0=71.54, t0=300e-15, sigma=50e-15, nu_L=0.29297e15)


def make_problem(user: dict, apply_auto: bool = False) -> Problem:
    """Config dict (same keys as the reference JSON) -> Problem.

    Restates config.py:83-311 (defaults), task_manager.py:350-517 (setup),
    task_manager.py:541-948 (potentials, initial/goal states, factory) and, when
    ``apply_auto`` is set, the *_auto rules of newcheb.py:913-936 (the
    reference's functional tests bypass newcheb and therefore do NOT apply
    sigma/t0/nt auto rules; nu_L_auto is applied by the task manager itself).
    """
    c = dict(_DEFAULTS)
    c.update({k: v for k, v in user.items() if k != "fitter"})
    f = dict(_FIT_DEFAULTS)
    uf = user.get("fitter", {})
    f.update({k: v for k, v in uf.items() if k != "propagation"})
    p = dict(_PROP_DEFAULTS)
    p.update(uf.get("propagation", {}))
    for d in (c, f, p):
        for k, v in d.items():
            if isinstance(v, float):
                d[k] = np.float64(v)
    tt = c["task_type"].lower()
    n = int(c["np"])
    nb, nlevs = int(c["nb"]), int(c["nlevs"])
    T = c["T"]
    if apply_auto:
        ig = c["init_guess"].lower()
        if c["sigma_auto"]:
            p["sigma"] = {"sqrsin": np.float64(2.0 * T), "maxwell": np.float64(T / 5.0),
                          "gauss": np.float64(T / 8.0)}.get(ig, p["sigma"])
        if c["t0_auto"]:
            p["t0"] = {"sqrsin": np.float64(0.0), "maxwell": np.float64(0.0),
                       "gauss": np.float64(T / 2.0)}.get(ig, p["t0"])
        if c["nt_auto"]:
            p["nt"] = math.floor(T / 2.0E-15)
    ht = c["hamil_type"].lower()
    if ht == "ntriv":
        ntriv, hamil = 1, "ntriv"
    elif ht == "bh_model":
        if c["lf_aug_type"].lower() == "z":
            ntriv, hamil = -1, "bh_z"
        else:
            ntriv, hamil = -2, "bh_x"
    elif ht == "two_levels":
        ntriv, hamil = 0, "two_levels"
    else:
        raise RuntimeError("Impossible case in the HamilType class")
    hamil = user.get("_hamil_override", hamil)
    ntriv = user.get("_ntriv_override", ntriv)
    F_goal = {"sm": -nb * nb, "re": -nb, "ss": -nb}[f["F_type"].lower()]
    dx, x = grid(c["L"], n)
    m = p["m"]
    wf = c["wf_type"].lower()

    def wf_init(x0, De, a):
        if wf == "morse":
            return wf_morse(x, x0, m, De, a)
        if wf == "harmonic":
            return wf_harmonic(x, x0, c["p0"], a)
        return wf_one(x, c["L"])

    zeros = lambda: np.zeros(n, np.complex128)
    pot = c["pot_type"].lower()
    single = tt in ("filtering", "single_pot")
    ut_manager = tt == "optimal_control_unit_transform" or (not single and pot == "none")
    psi0 = np.zeros((nb, nlevs, n), np.complex128)
    psif = np.zeros((nb, nlevs, n), np.complex128)
    v = np.zeros((nlevs, n), np.float64)
    vmin = np.zeros(nlevs, np.float64)
    if ut_manager:                                        # task_manager.py:748-892
        if nb == nlevs:
            for b in range(nb):
                psi0[b, b] = wf_init(c["x0"], c["De"], c["a"])
            om = cmath.exp(1j * 2.0 * math.pi / nb)
            Fm = np.zeros((nb, nb), np.complex128)
            for a_ in range(nb):
                for b_ in range(nb):
                    Fm[a_, b_] = om ** (a_ * b_) / math.sqrt(nb)
            for b in range(nb):
                for g in range(nb):
                    acc = np.zeros(n, np.complex128)
                    for i in range(nb):
                        acc = np.add(acc, Fm[g, i] * psi0[b, i])
                    psif[b, g] = acc
        elif nb == 1 and nlevs == 2:
            psi0[0, 0] = wf_init(c["x0"], c["De"], c["a"])
            psif[0, 0] = wf_init(c["x0"], c["De"], c["a"]) / math.sqrt(2.0)
            psif[0, 1] = wf_init(c["x0"], c["De"], c["a"]) / math.sqrt(2.0)
        else:
            raise RuntimeError("Impossible Task Type")
        if not ntriv:
            D_l = -c["Du"] / 2.0
            v[0, :] = D_l
            vmin[0] = D_l
            v[1, :] = c["Du"] + D_l
            vmin[1] = c["Du"] + D_l
        elif ntriv < 0:
            U, W = c["U"], c["W"]
            Emax = p["E0"] * f["Em"]
            l = (nlevs - 1) / 2.0
            if ntriv == -1:
                vmx = 2.0 * U * l ** 2 + 2.0 * Emax * l
                vmn = -2.0 * Emax * l
            else:
                vmx = 2.0 * l * (U + W * l)
                if W != 0.0 and nlevs >= U / W + 1:
                    vmn = -U * U / W / 2.0
                else:
                    vmn = 2.0 * l * (-U + W * l)
            v[:, :] = vmx
            vmin[:] = vmn
        else:
            raise RuntimeError("Unsupported type of potential!")
    else:
        assert nlevs == 2
        if pot == "morse":                                # task_manager.py:653-745
            v[0] = np.array([c["De"] * (1.0 - math.exp(-c["a"] * xi)) * (1.0 - math.exp(-c["a"] * xi)) for xi in x])
            if not single:
                v[1] = np.array([c["De_e"] * (1.0 - math.exp(-c["a_e"] * (xi - c["x0p"]))) *
                                 (1.0 - math.exp(-c["a_e"] * (xi - c["x0p"]))) + c["Du"] for xi in x])
                vmin[1] = c["Du"]
        elif pot == "harmonic":                           # task_manager.py:541-650
            k_s = HART_TO_CM / m / DALT_TO_AU / pow(c["a"], 4.0)
            v[0] = np.array([k_s * xi * xi / 2.0 for xi in x])
            if not single:
                k_u = HART_TO_CM / m / DALT_TO_AU / pow(c["a_e"], 4.0)
                v[1] = np.array([k_u * (xi - c["x0p"]) * (xi - c["x0p"]) / 2.0 + c["Du"] for xi in x])
                vmin[1] = c["Du"]
        else:
            raise RuntimeError("Impossible PotentialType")
        for b in range(nb):
            psi0[b, 0] = wf_init(c["x0"], c["De"], c["a"])
            if single:
                psif[b, 0] = wf_harmonic(x, c["x0"], c["p0"], c["a"]) if pot == "harmonic" \
                    else wf_morse(x, c["x0"], m, c["De"], c["a"])
            else:
                psif[b, 1] = wf_harmonic(x, c["x0p"] + c["x0"], c["p0"], c["a_e"]) if pot == "harmonic" \
                    else wf_morse(x, c["x0p"] + c["x0"], m, c["De_e"], c["a_e"])
    akx2 = initak(n, dx, 2, ntriv)
    akx2 *= -HART_TO_CM / (2.0 * m * DALT_TO_AU)
    nt = int(p["nt"])
    t_step = T / nt
    t_list = [t_step * l for l in range(nt + 1)]
    nu = np.float64(1.0 / 2.0 / T) if c["nu_L_auto"] else p["nu_L"]
    init_dir = BACKWARD if tt == "optimal_control_unit_transform" else FORWARD
    return Problem(
        task_type=tt, hamil=hamil, ntriv=ntriv, np_=n, nlevs=nlevs, nb=nb, L=c["L"], T=T, dx=dx, x=x,
        v=v, vmin=vmin, akx2=akx2, psi0=psi0, psif=psif, U=c["U"], W=c["W"], delta=c["delta"], m=m,
        nch=int(p["nch"]), nt=nt, E0=p["E0"], t0=p["t0"], sigma=p["sigma"], nu_L=p["nu_L"], nu=nu,
        envelope=c["init_guess"].lower(), carrier=c["init_guess_hf"].lower(), pcos=f["pcos"],
        w_list=list(f["w_list"]), hf_hide=bool(f["hf_hide"]), h_lambda=f["h_lambda"],
        h_lambda_mode=f["h_lambda_mode"].lower(), epsilon=f["epsilon"], iter_max=int(f["iter_max"]),
        iter_mid_1=int(f["iter_mid_1"]), iter_mid_2=int(f["iter_mid_2"]), q=f["q"],
        impulses_number=int(f["impulses_number"]), delay=f["delay"], k_E=f["k_E"], lamb=f["lamb"],
        pow_=f["pow"], F_type=f["F_type"].lower(),
        F_goal=F_goal, init_dir=init_dir, t_step=t_step, t_list=t_list)
