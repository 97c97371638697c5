"""Task set-up: initial / goal states, potentials, kinetic multiplier, laser envelope and carrier,
functionals -- what the reference's ``task_manager.create(conf)`` hands to ``FittingSolver``
(reference task_manager.py:18-948), with the same attribute names on the returned object:

    psi0, psif (PsiBasis)   v = [(vmin_n, v_n[np]) ...]   akx2   hamil2D   ntriv   init_dir
    dx, x, t_step, t_list   F_goal, F_type, aF_type       laser_field, laser_field_hf, nu

This runs once per run and is O(np) host arithmetic; it is evaluated with Python ``math`` scalars
element by element, like the reference, because these arrays are INPUTS of the GPU path and have
to be bit-identical for the parity tests to mean anything.
"""
from __future__ import annotations

import cmath
import copy
import math
import threading

import numpy

from . import functionals, grid_setup, hamil_2d, math_base
from .config import TaskRootConfiguration
from .phys_consts import DALT_TO_AU, HART_TO_CM
from .psi_basis import Psi, PsiBasis

dalt_to_au = DALT_TO_AU
hart_to_cm = HART_TO_CM

_T = TaskRootConfiguration
_F = TaskRootConfiguration.FitterConfiguration


class Direction:
    """Stand-in with the two members of PropagationSolver.Direction (propagation.py:102-104); the real
    enum lives in qcontrol_b200.propagation and compares equal by ``.value``."""
    FORWARD = 1
    BACKWARD = -1


# ----------------------------------------------------------------------------- envelopes (task_manager.py:18-70)
class _LaserFields:
    @staticmethod
    def zero(E0, t, t0, sigma):
        return numpy.float64(0.0)

    @staticmethod
    def const(E0, t, t0, sigma):
        return E0

    @staticmethod
    def laser_field_gauss(E0, t, t0, sigma):
        d = t - t0
        return E0 * math.exp(-d * d / 2.0 / sigma / sigma)

    @staticmethod
    def laser_field_sqrsin(E0, t, t0, sigma):
        s = math.sin(2.0 * math.pi * (t - t0) / sigma)
        return E0 * s * s

    @staticmethod
    def laser_field_maxwell(E0, t, t0, sigma):
        d = t - t0
        return E0 * t * t * math.exp(-d * d / 2.0 / sigma / sigma) / sigma / sigma


# ----------------------------------------------------------------------------- carriers (task_manager.py:72-159)
class _LaserFieldsHighFrequencyPart:
    @staticmethod
    def cexp(nu, freq_mult, t, pcos, w_list):
        return numpy.complex128(cmath.exp(1j * 2.0 * math.pi * nu * freq_mult * t))

    @staticmethod
    def cos(nu, freq_mult, t, pcos, w_list):
        return numpy.float64(math.cos(2.0 * math.pi * pcos * nu * freq_mult * t))

    @staticmethod
    def sin(nu, freq_mult, t, pcos, w_list):
        return numpy.float64(math.sin(2.0 * math.pi * pcos * nu * freq_mult * t))

    @staticmethod
    def cos_set(nu, freq_mult, t, pcos, w_list):
        base = 2.0 * math.pi * nu * freq_mult * t
        acc = w_list[0] * math.cos(base)
        k = -1
        for p in range(2, math.floor(pcos) + 1):
            k += 2
            acc += w_list[k] * math.cos(base * p)
            acc += w_list[k + 1] * math.cos(base / p)
        return numpy.float64(acc)

    @staticmethod
    def sin_set(nu, freq_mult, t, pcos, w_list):
        base = 2.0 * math.pi * nu * freq_mult * t
        acc = w_list[0] * math.sin(base)
        for p in range(1, math.floor(pcos) + 1):
            acc += w_list[p] * math.sin(base * (2 * p + 1))
        return numpy.float64(acc)


# ----------------------------------------------------------------------------- functionals (task_manager.py:162-268)
def _tagged(kind, fn):
    fn.kind = kind
    return staticmethod(fn)


class _F_type:
    """F_type(gc_vec, nb) callables with the reference's names; the arithmetic lives in ``functionals``."""
    F_sm = _tagged("sm", lambda gc_vec, nb: functionals.F_value("sm", gc_vec, nb))
    F_re = _tagged("re", lambda gc_vec, nb: functionals.F_value("re", gc_vec, nb))
    F_ss = _tagged("ss", lambda gc_vec, nb: functionals.F_value("ss", gc_vec, nb))


def _a_of(kind):
    def a(psi_init, chi_init, dx, nb, nlevs, np):
        assert len(psi_init) == nb and len(psi_init.psis[0].f) == nlevs and psi_init.psis[0].f[0].size == np
        return functionals.a_value(kind, psi_init.as_array(), chi_init.as_array(), dx)
    return _tagged(kind, a)


class _aF_type:
    """aF_type(psi_init, chi_init, dx, nb, nlevs, np) callables with the reference's names."""
    a_sm = _a_of("sm")
    a_re = _a_of("re")
    a_ss = _a_of("ss")


# ----------------------------------------------------------------------------- wavefunctions (task_manager.py:279-337)
class _PsiFunctions:
    @staticmethod
    def zero(np):
        return numpy.zeros(np, dtype=numpy.complex128)

    @staticmethod
    def one(x, np, x0, p0, m, De, a, L):
        return numpy.array([1.0 / math.sqrt(L)] * np).astype(numpy.complex128)

    @staticmethod
    def harmonic(x, np, x0, p0, m, De, a, L):
        norm4 = pow(math.pi, 0.25)
        root = math.sqrt(a)
        return numpy.array([cmath.exp(-(xi - x0) * (xi - x0) / 2.0 / a / a + 1j * p0 * xi) / norm4 / root
                            for xi in x]).astype(numpy.complex128)

    @staticmethod
    def morse(x, np, x0, p0, m, De, a, L):
        omega_0 = a * math.sqrt(2.0 * De / hart_to_cm / m / dalt_to_au) * hart_to_cm
        xe = omega_0 / 4.0 / De
        arg = 1.0 / xe - 1.0
        pref = math.sqrt(a / math.gamma(arg))
        half = numpy.float64(arg / 2.0)
        out = []
        for xi in x:
            yi = math.exp(-a * (xi - x0)) / xe
            out.append(pref * math.exp(-yi / 2.0) * pow(yi, half))
        return numpy.array(out).astype(numpy.complex128)


# ----------------------------------------------------------------------------- potentials
def _harmonic_lower(x, m, a):
    """task_manager.py:546-565"""
    k_s = hart_to_cm / m / dalt_to_au / pow(a, 4.0)
    return [(numpy.float64(0.0), numpy.array([k_s * xi * xi / 2.0 for xi in x]))]


def _morse_lower(x, m, De, a):
    """task_manager.py:657-676"""
    return [(numpy.float64(0.0),
             numpy.array([De * (1.0 - math.exp(-a * xi)) * (1.0 - math.exp(-a * xi)) for xi in x]))]


def _flat_upper(np):
    return (numpy.float64(0.0), numpy.array([0.0] * np).astype(numpy.float64))


def _level_window(x, np, ntriv, conf_task, Du):
    """Constant 'potentials' of grid-free systems; only their extremes matter (task_manager.py:824-870)."""
    v = []
    if not ntriv:
        D_l = -Du / 2.0
        v.append((D_l, numpy.array([D_l] * np)))
        D_u = Du + D_l
        v.append((D_u, numpy.array([D_u] * np)))
        return v
    if ntriv > 0:
        raise RuntimeError("Unsupported type of potential!")
    U, W = conf_task.U, conf_task.W
    Emax = conf_task.fitter.propagation.E0 * conf_task.fitter.Em
    l = (conf_task.nlevs - 1) / 2.0
    if ntriv == -1:
        vmax = 2.0 * U * l ** 2 + 2.0 * Emax * l
        vmin = -2.0 * Emax * l
    elif ntriv == -2:
        vmax = 2.0 * l * (U + W * l)
        if W != 0.0 and conf_task.nlevs >= U / W + 1:
            vmin = -U * U / W / 2.0
        else:
            vmin = 2.0 * l * (-U + W * l)
    else:
        raise RuntimeError("Impossible case in the LfAugType class")
    for _ in range(conf_task.nlevs):
        v.append((vmin, numpy.array([vmax] * np)))
    return v


_ENVELOPES = {_F.InitGuess.ZERO: _LaserFields.zero, _F.InitGuess.CONST: _LaserFields.const,
              _F.InitGuess.GAUSS: _LaserFields.laser_field_gauss, _F.InitGuess.SQRSIN: _LaserFields.laser_field_sqrsin,
              _F.InitGuess.MAXWELL: _LaserFields.laser_field_maxwell}
_CARRIERS = {_F.InitGuessHf.EXP: _LaserFieldsHighFrequencyPart.cexp, _F.InitGuessHf.COS: _LaserFieldsHighFrequencyPart.cos,
             _F.InitGuessHf.SIN: _LaserFieldsHighFrequencyPart.sin,
             _F.InitGuessHf.COS_SET: _LaserFieldsHighFrequencyPart.cos_set,
             _F.InitGuessHf.SIN_SET: _LaserFieldsHighFrequencyPart.sin_set}
_WAVEFUNCS = {_T.WaveFuncType.MORSE: _PsiFunctions.morse, _T.WaveFuncType.HARMONIC: _PsiFunctions.harmonic,
              _T.WaveFuncType.CONST: _PsiFunctions.one}
_FUNCTIONALS = {_F.FType.SM: (_F_type.F_sm, _aF_type.a_sm, lambda nb: -nb * nb),
                _F.FType.RE: (_F_type.F_re, _aF_type.a_re, lambda nb: -nb),
                _F.FType.SS: (_F_type.F_ss, _aF_type.a_ss, lambda nb: -nb)}


# The O(np) arrays of a run (grid, initial / goal states, potentials, kinetic multiplier) are evaluated element by
# element with Python scalars so that they are bit-identical to the reference's -- 5 ms per run at np = 2048.  The
# variants of a batch template differ in one or two scalars (run id, T, E0, ...; input_batch/*.json), so most of
# them ask for arrays that were just built: results are kept per argument tuple and handed out as deep copies
# (every run still owns its arrays, like in the reference).  Bounded, per thread (the reference farms jobs over
# threads, newcheb.py:656-659).
_MEMO = threading.local()
_MEMO_MAX = 64


def _memo(kind, key, build):
    store = getattr(_MEMO, "store", None)
    if store is None:
        store = _MEMO.store = {}
    k = (kind,) + tuple(key)
    if k not in store:
        if len(store) >= _MEMO_MAX:
            store.clear()
        store[k] = build()
    return copy.deepcopy(store[k])


class TaskManager:
    """Common set-up (task_manager.py:350-517); subclasses choose states and potentials."""
    family = None       # 'single' | 'multiple' | 'unit_transform'
    shape = None        # 'morse' | 'harmonic' | None

    def __init__(self, conf_task: TaskRootConfiguration):
        c = conf_task
        self.conf_task = c
        self.psi_init_impl = _WAVEFUNCS[c.wf_type]
        self.lf_init_guess = _ENVELOPES[c.init_guess]
        if c.init_guess_hf in (_F.InitGuessHf.COS_SET, _F.InitGuessHf.SIN_SET) and not c.fitter.pcos > 1.0:
            raise ValueError("The maximum frequency multiplier in the high frequency part of the laser field, "
                             "'pcos', has to be > 1")
        self.lf_init_guess_hf = _CARRIERS[c.init_guess_hf]
        if c.hamil_type == _T.HamilType.NTRIV:
            self.ntriv, self.hamil_impl = 1, hamil_2d.Hamil2DNonTrivial
            if c.init_guess_hf != _F.InitGuessHf.EXP:
                raise RuntimeError("For a non-trivial Hamiltonian an exponential high-frequency part of initial "
                                   "guess for the laser field has to be used!")
        elif c.hamil_type == _T.HamilType.BH_MODEL:
            if c.lf_aug_type == _F.LfAugType.Z:
                self.ntriv, self.hamil_impl = -1, hamil_2d.Hamil2DBHZ
            elif c.lf_aug_type == _F.LfAugType.X:
                self.ntriv, self.hamil_impl = -2, hamil_2d.Hamil2DBHX
            else:
                raise RuntimeError("Impossible case in the LfAugType class")
        elif c.hamil_type == _T.HamilType.TWO_LEVELS:
            self.ntriv, self.hamil_impl = 0, hamil_2d.Hamil2DTwoLevels
            if c.nb != 2:
                raise RuntimeError("Number of basis vectors 'nb' for 'hamil_type' = 'two_levels' has to be equal to 2!")
            if c.fitter.hf_hide:
                raise RuntimeError("'hf_hide' parameter for 'hamil_type' = 'two_levels' should be set to 'false'!")
        else:
            raise RuntimeError("Impossible case in the HamilType class")
        F_fn, a_fn, goal = _FUNCTIONALS[c.fitter.F_type]
        self._F_fn, self._a_fn, self.F_goal = F_fn, a_fn, goal(c.nb)
        self.init_dir = Direction.FORWARD
        self.dx, self.x = _memo("grid", (c.np, float(c.L)), lambda: grid_setup.GridConstructor(c).grid_setup())
        cp = c.fitter.propagation
        args = (self.x, c.np, c.x0, c.p0, c.x0p, cp.m, c.De, c.De_e, c.Du, c.a, c.a_e, c.L, c.nb, c.nlevs)
        who = (type(self).__name__, c.wf_type, c.hamil_type, c.lf_aug_type) + tuple(float(a) for a in args[1:])
        self.psi0 = _memo("psi0", who, lambda: self.psi_init(*args))
        self.psif = _memo("psif", who, lambda: self.psi_goal(*args))
        self.t_step, self.t_list = grid_setup.ForwardTimeGridConstructor(conf_task=c).grid_setup()
        self.v = _memo("pot", who + (float(c.U), float(c.W), float(cp.E0), float(c.fitter.Em)),
                       lambda: self.pot(self.x, c.np, cp.m, c.De, c.a, c.x0p, c.De_e, c.a_e, c.Du))

        def kinetic():
            ak = math_base.initak(c.np, self.dx, 2, self.ntriv)
            ak *= -hart_to_cm / (2.0 * cp.m * dalt_to_au)
            return ak
        self.akx2 = _memo("akx2", (c.np, float(self.dx), self.ntriv, float(cp.m)), kinetic)
        self.hamil2D = self.hamil_impl(self.v, self.akx2, c.np, c.U, c.W, c.delta, self.ntriv)
        self.nu = numpy.float64(1.0 / 2.0 / c.T) if c.nu_L_auto else cp.nu_L

    # -- callables handed to FittingSolver ------------------------------------------------------
    def laser_field(self, E0, t, t0, sigma):
        return self.lf_init_guess(E0, t, t0, sigma)

    def laser_field_hf(self, freq_mult, t, pcos, w_list):
        return self.lf_init_guess_hf(self.nu, freq_mult, t, pcos, w_list)

    def F_type(self, gc_vec, nb):
        return self._F_fn(gc_vec, nb)

    def aF_type(self, psi_init, chi_init, dx, nb, nlevs, np):
        return self._a_fn(psi_init, chi_init, dx, nb, nlevs, np)

    # -- states ------------------------------------------------------------------------------------
    def _goal_fn(self):
        return _PsiFunctions.harmonic if self.shape == "harmonic" else _PsiFunctions.morse

    def psi_init(self, x, np, x0, p0, x0p, m, De, De_e, Du, a, a_e, L, nb, nlevs) -> PsiBasis:
        out = PsiBasis(nb)
        for b in range(nb):
            out.psis[b].f[0] = self.psi_init_impl(x, np, x0, p0, m, De, a, L)
            out.psis[b].f[1] = _PsiFunctions.zero(np)
        return out

    def psi_goal(self, x, np, x0, p0, x0p, m, De, De_e, Du, a, a_e, L, nb, nlevs) -> PsiBasis:
        out = PsiBasis(nb)
        goal = self._goal_fn()
        for b in range(nb):
            if self.family == "single":       # stay on the lower state (task_manager.py:596-601, 711-716)
                out.psis[b].f[0] = goal(x, np, x0, p0, m, De, a, L)
                out.psis[b].f[1] = _PsiFunctions.zero(np)
            else:                             # land on the shifted upper state (task_manager.py:639-644, 757-762)
                out.psis[b].f[0] = _PsiFunctions.zero(np)
                out.psis[b].f[1] = goal(x, np, x0p + x0, p0, m, De_e, a_e, L)
        return out

    def pot(self, x, np, m, De, a, x0p, De_e, a_e, Du):
        raise NotImplementedError()


class HarmonicSingleStateTaskManager(TaskManager):
    family, shape = "single", "harmonic"

    @staticmethod
    def _pot_level1(x, m, a):
        return _harmonic_lower(x, m, a)

    def pot(self, x, np, m, De, a, x0p, De_e, a_e, Du):
        return _harmonic_lower(x, m, a) + [_flat_upper(np)]


class HarmonicMultipleStateTaskManager(HarmonicSingleStateTaskManager):
    family = "multiple"

    def pot(self, x, np, m, De, a, x0p, De_e, a_e, Du):
        v = _harmonic_lower(x, m, a)
        k_u = hart_to_cm / m / dalt_to_au / pow(a_e, 4.0)
        v.append((Du, numpy.array([k_u * (xi - x0p) * (xi - x0p) / 2.0 + Du for xi in x])))
        return v


class MorseSingleStateTaskManager(TaskManager):
    family, shape = "single", "morse"

    @staticmethod
    def _pot_level1(x, m, De, a):
        return _morse_lower(x, m, De, a)

    def pot(self, x, np, m, De, a, x0p, De_e, a_e, Du):
        return _morse_lower(x, m, De, a) + [_flat_upper(np)]


class MorseMultipleStateTaskManager(MorseSingleStateTaskManager):
    family = "multiple"

    @staticmethod
    def _pot(x, np, m, De, a, x0p, De_e, a_e, Du, ntriv, conf_task):
        v = _morse_lower(x, m, De, a)
        v.append((Du, numpy.array([De_e * (1.0 - math.exp(-a_e * (xi - x0p))) * (1.0 - math.exp(-a_e * (xi - x0p))) + Du
                                   for xi in x])))
        return v

    def pot(self, x, np, m, De, a, x0p, De_e, a_e, Du):
        return self._pot(x, np, m, De, a, x0p, De_e, a_e, Du, self.ntriv, self.conf_task)


class MultipleStateUnitTransformTaskManager(MorseMultipleStateTaskManager):
    family = "unit_transform"

    @staticmethod
    def _quant_fourier_transform(nb):
        """Goal matrix of the quantum Fourier transform (task_manager.py:756-764)."""
        omega = cmath.exp(1j * 2.0 * math.pi / nb)
        F = numpy.zeros((nb, nb), dtype=numpy.complex128)
        for r in range(nb):
            for c in range(nb):
                F[r, c] = omega ** (r * c) / math.sqrt(nb)
        return F

    @staticmethod
    def _matrix_PsiBasis_mult(F, psi: PsiBasis, nb, np):
        out = PsiBasis(nb, nb)
        for b in range(nb):
            for g in range(nb):
                acc = numpy.zeros(np, dtype=numpy.complex128)
                for i in range(nb):
                    acc = numpy.add(acc, complex(F[g, i]) * psi.psis[b].f[i])
                out.psis[b].f[g] = acc
        return out

    def psi_init(self, x, np, x0, p0, x0p, m, De, De_e, Du, a, a_e, L, nb, nlevs) -> PsiBasis:
        if nb == nlevs:
            out = PsiBasis(nb, nb)
            for b in range(nb):
                for n in range(nb):
                    out.psis[b].f[n] = self.psi_init_impl(x, np, x0, p0, m, De, a, L) if n == b else _PsiFunctions.zero(np)
            return out
        if nb == 1 and nlevs == 2:
            out = PsiBasis(nb)
            out.psis[0].f[0] = self.psi_init_impl(x, np, x0, p0, m, De, a, L)
            out.psis[0].f[1] = _PsiFunctions.zero(np)
            return out
        raise RuntimeError("Impossible Task Type")

    def psi_goal(self, x, np, x0, p0, x0p, m, De, De_e, Du, a, a_e, L, nb, nlevs) -> PsiBasis:
        if nb == nlevs:
            basis = self.psi_init(x, np, x0, p0, x0p, m, De, De_e, Du, a, a_e, L, nb, nlevs)
            return self._matrix_PsiBasis_mult(self._quant_fourier_transform(nb), basis, nb, np)
        if nb == 1 and nlevs == 2:
            out = PsiBasis(nb)
            out.psis[0].f[0] = self.psi_init_impl(x, np, x0, p0, m, De, a, L) / math.sqrt(2.0)
            out.psis[0].f[1] = self.psi_init_impl(x, np, x0, p0, m, De, a, L) / math.sqrt(2.0)
            return out
        raise RuntimeError("Impossible Task Type")

    @staticmethod
    def _pot(x, np, m, De, a, x0p, De_e, a_e, Du, ntriv, conf_task):
        return _level_window(x, np, ntriv, conf_task, Du)


def create(conf_task: TaskRootConfiguration):
    """Factory with the reference's validation rules (task_manager.py:895-948)."""
    c = conf_task
    if c.task_type in (_T.TaskType.FILTERING, _T.TaskType.SINGLE_POT):
        if c.nlevs != 2:
            raise RuntimeError("Number of levels in basis vectors 'nlevs' for 'task_type' = 'single_pot' and "
                               "'filtering' should be equal to 2!")
        if c.hamil_type != _T.HamilType.NTRIV:
            raise RuntimeError("For 'task_type' = 'single_pot' and 'filtering' the Hamiltonian type 'hamil_type' = "
                               "'ntriv' should be specified!")
        if c.nb != 1:
            raise RuntimeError("Number of basis vectors 'nb' for 'task_type' = 'single_pot' and 'filtering' should "
                               "be equal to 1!")
        if c.pot_type == _T.PotentialType.MORSE:
            return MorseSingleStateTaskManager(c)
        if c.pot_type == _T.PotentialType.HARMONIC:
            return HarmonicSingleStateTaskManager(c)
        raise RuntimeError("Impossible PotentialType")
    if c.task_type == _T.TaskType.OPTIMAL_CONTROL_UNIT_TRANSFORM:
        if c.pot_type != _T.PotentialType.NONE:
            raise RuntimeError("Impossible PotentialType for the unitary transformation task")
        if c.fitter.hf_hide:
            raise RuntimeError("'hf_hide' parameter for 'none' potential type should be set to 'false'!")
        tm = MultipleStateUnitTransformTaskManager(c)
        tm.init_dir = Direction.BACKWARD
        return tm
    if c.nb != 1:
        raise RuntimeError("Number of basis vectors 'nb' for 'task_type' = 'trans_wo_control', 'intuitive_control', "
                           "'local_control_population', 'local_control_projection', and 'optimal_control_krotov' "
                           "should be equal to 1!")
    if c.nlevs != 2:
        raise RuntimeError("Number of levels in basis vectors 'nlevs' for 'task_type' = 'trans_wo_control', "
                           "'intuitive_control', 'local_control_population', 'local_control_projection', and "
                           "'optimal_control_krotov' should be equal to 2!")
    if c.pot_type == _T.PotentialType.MORSE:
        return MorseMultipleStateTaskManager(c)
    if c.pot_type == _T.PotentialType.HARMONIC:
        return HarmonicMultipleStateTaskManager(c)
    if c.pot_type == _T.PotentialType.NONE:
        if c.fitter.hf_hide:
            raise RuntimeError("'hf_hide' parameter for 'none' potential type should be set to 'false'!")
        return MultipleStateUnitTransformTaskManager(c)
    raise RuntimeError("Impossible PotentialType")
