"""Synthetic batches of independent runs (BASELINE.json config 5).

The reference's batch inputs (input_batch/*.json) vary one or two scalars per run; the synthetic
Krotov batch does the same at scale: the two-state Morse problem of input_files/input_task_krotov.json
re-gridded to ``np_`` points, with the excited-state well depth, its displacement and the pulse
amplitude jittered per run from a seeded generator.  Arrays are built with
vectorised numpy -- these runs have no reference counterpart to be bit-identical to; parity of the
same code path is established on the reference's own configurations in tests/.
"""
from __future__ import annotations

import cmath
import math
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

from .phys_consts import DALT_TO_AU, HART_TO_CM


@dataclass
class RunProblem:
    """What one run hands to the control loops (the products of TaskManager, task_manager.py:350-517)."""
    task_type: str
    hamil: str
    ntriv: int
    np_: int
    nlevs: int
    nb: int
    L: float
    T: float
    dx: float
    x: np.ndarray
    v: np.ndarray
    vmin: np.ndarray
    akx2: np.ndarray
    psi0: np.ndarray
    psif: np.ndarray
    m: float
    nch: int
    nt: int
    E0: float
    t0: float
    sigma: float
    nu_L: float
    nu: float
    hf_hide: bool
    h_lambda: float
    t_step: float
    t_list: List[float]
    U: float = 0.0
    W: float = 0.0
    delta: float = 1.0
    pcos: float = 1.0
    w_list: Sequence[float] = ()
    h_lambda_mode: str = "const"
    epsilon: float = 1e-8
    iter_max: int = 1
    iter_mid_1: int = 0
    iter_mid_2: int = 1
    q: float = 0.0
    impulses_number: int = 1
    delay: float = 0.0
    F_type: str = "sm"
    F_goal: float = -1.0
    envelope: str = "gauss"
    carrier: str = "exp"

    def laser_field(self, E0, t, t0, sigma):            # gaussian envelope, task_manager.py:27-40
        return E0 * math.exp(-(t - t0) * (t - t0) / 2.0 / sigma / sigma)

    def laser_field_hf(self, freq_mult, t, pcos, w_list):   # exp carrier, task_manager.py:73-87
        return np.complex128(cmath.exp(1j * 2.0 * math.pi * self.nu * freq_mult * t))


def _morse_ground(x, x0, m, De, a):
    """Morse ground state (task_manager.py:311-337), vectorised."""
    omega_0 = a * math.sqrt(2.0 * De / HART_TO_CM / m / DALT_TO_AU) * HART_TO_CM
    xe = omega_0 / 4.0 / De
    arg = 1.0 / xe - 1.0
    y = np.exp(-a * (x - x0)) / xe
    lg = 0.5 * (math.log(a) - math.lgamma(arg)) - y / 2.0 + (arg / 2.0) * np.log(y)
    return np.exp(lg).astype(np.complex128)


def krotov_batch(batch: int, np_: int = 2048, nt: int = 16, nch: int = 64, seed: int = 12345,
                 jitter: bool = True) -> List[RunProblem]:
    """``batch`` independent two-level Krotov problems on an ``np_``-point grid, ``nt`` time steps of the
    reference's step size (350 fs / 230000), pulse centred in the window."""
    rng = np.random.default_rng(seed)
    L, m, a = 5.0, 0.5, 1.0
    dt = 350e-15 / 230000
    T = dt * nt
    dx = L / (np_ - 1)
    x = (np.arange(np_, dtype=np.float64) - (np_ - 1) / 2.0) * dx
    dk = 2.0 * math.pi / (np_ - 1) / dx
    k = np.zeros(np_)
    idx = np.arange(1, np_ // 2 + 1)
    k[idx] = idx
    k[np_ - idx] = idx
    akx2 = (-(dk * k) ** 2 * (-HART_TO_CM / (2.0 * m * DALT_TO_AU))).astype(np.complex128)
    v_l = 20000.0 * (1.0 - np.exp(-a * x)) ** 2
    psi_g = _morse_ground(x, 0.0, m, 20000.0, a)
    t_step = T / nt
    t_list = [t_step * l for l in range(nt + 1)]
    out = []
    for b in range(batch):
        u = rng.uniform(-1.0, 1.0, 3) if jitter else np.zeros(3)
        De_e = 10000.0 * (1.0 + 0.05 * u[0])
        x0p = -0.17 + 0.02 * u[1]
        E0 = 2861.6 * (1.0 + 0.2 * u[2])      # 40 x the reference amplitude: the short window still transfers population
        Du = 20000.0
        v_u = De_e * (1.0 - np.exp(-a * (x - x0p))) ** 2 + Du
        psi0 = np.zeros((1, 2, np_), np.complex128)
        psi0[0, 0] = psi_g
        psif = np.zeros((1, 2, np_), np.complex128)
        psif[0, 1] = _morse_ground(x, x0p, m, De_e, a)
        out.append(RunProblem(
            task_type="optimal_control_krotov", hamil="ntriv", ntriv=1, np_=np_, nlevs=2, nb=1, L=L, T=T, dx=dx, x=x,
            v=np.stack([v_l, v_u]), vmin=np.array([0.0, Du]), akx2=akx2, psi0=psi0, psif=psif, m=m, nch=nch, nt=nt,
            E0=np.float64(E0), t0=np.float64(T / 2), sigma=np.float64(T / 6), nu_L=np.float64(0.29297e15),
            nu=np.float64(0.29297e15), hf_hide=True, h_lambda=np.float64(0.0066), t_step=t_step, t_list=t_list))
    return out
