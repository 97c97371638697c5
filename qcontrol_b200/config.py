"""Task configuration with the reference's JSON keys, defaults and attribute access
(reference config.py:10-311; the report-side configurations are I/O and out of scope).

``TaskRootConfiguration().load(dict)`` accepts exactly the dictionaries of
``input_files/*.json`` / ``tests/functional_tests.py``; values are reached as attributes
(``conf.fitter.propagation.nt``), floats become ``numpy.float64``, enum keys accept the lower-case
names used in the JSON files or their integer codes, unknown keys are ignored and missing keys
keep their defaults, as in the reference (config.py:27-64).
"""
from __future__ import annotations

import contextlib
import copy
import threading
from enum import Enum

import numpy

_QUIET = threading.local()


@contextlib.contextmanager
def quiet():
    """Silences the "parameter hasn't been provided" notes of load() for the calling THREAD (contextlib.redirect_stdout
    swaps sys.stdout for every thread; batch set-up runs beside the control loops of the previous sub-batch)."""
    prev = getattr(_QUIET, "on", False)
    _QUIET.on = True
    try:
        yield
    finally:
        _QUIET.on = prev


class _NamedEnum(Enum):
    @classmethod
    def from_str(cls, s):
        return cls[s.upper()]

    @classmethod
    def from_int(cls, i):
        return cls(i)


class ConfigurationBase:
    """Typed key/value store: the type of a key's default decides how a user value is converted."""

    def __init__(self, key_prefix: str = ""):
        self._empty = True
        self._key_prefix = key_prefix
        self._data = {}

    def _define(self, **defaults):
        for k, v in defaults.items():
            self._data[k] = numpy.float64(v) if isinstance(v, float) else v

    def is_empty(self):
        return self._empty

    def load(self, user_data):
        self._empty = False
        for key, default in self._data.items():
            if self._key_prefix + key in user_data:
                raw = user_data[self._key_prefix + key]
            elif key in user_data:
                raw = user_data[key]
            else:
                if not getattr(_QUIET, "on", False):
                    print("Parameter '%s' hasn't been provided in the input json file. "
                          "It will be calculated automatically or the default value will be used: %s" % (key, str(default)))
                continue
            # bool is tested after int on purpose: the reference does the same (config.py:41-44), so a
            # JSON true/false stored under a bool default goes through int()
            if isinstance(default, str):
                self._data[key] = str(raw)
            elif isinstance(default, float):
                self._data[key] = numpy.float64(raw)
            elif isinstance(default, int):
                self._data[key] = int(raw)
            elif isinstance(default, Enum):
                if isinstance(raw, str):
                    self._data[key] = type(default).from_str(raw)
                else:
                    self._data[key] = type(default).from_int(int(raw))
            elif isinstance(default, list):
                assert isinstance(raw, list), "The value for %s has to be of type 'list'" % key
                self._data[key] = raw
            elif isinstance(default, ConfigurationBase):
                default.load(raw)

    def __getattr__(self, key):
        data = self.__dict__.get("_data")
        if data is not None and key in data:
            return data[key]
        raise AttributeError(key)

    def __setattr__(self, key, value):
        data = self.__dict__.get("_data")
        if data is not None and key in data:
            data[key] = value
        else:
            object.__setattr__(self, key, value)

    def __deepcopy__(self, memo):
        out = type(self).__new__(type(self))
        memo[id(self)] = out
        for k, v in self.__dict__.items():
            object.__setattr__(out, k, copy.deepcopy(v, memo))
        return out


class TaskRootConfiguration(ConfigurationBase):
    class TaskType(_NamedEnum):
        SINGLE_POT = 0
        FILTERING = 1
        TRANS_WO_CONTROL = 2
        INTUITIVE_CONTROL = 3
        LOCAL_CONTROL_POPULATION = 4
        LOCAL_CONTROL_PROJECTION = 5
        OPTIMAL_CONTROL_KROTOV = 6
        OPTIMAL_CONTROL_UNIT_TRANSFORM = 7

    class PotentialType(_NamedEnum):
        MORSE = 0
        HARMONIC = 1
        NONE = 2

    class WaveFuncType(_NamedEnum):
        MORSE = 0
        HARMONIC = 1
        CONST = 2

    class HamilType(_NamedEnum):
        NTRIV = 0
        TWO_LEVELS = 1
        BH_MODEL = 2

    class FitterConfiguration(ConfigurationBase):
        class InitGuess(_NamedEnum):
            ZERO = 0
            CONST = 1
            GAUSS = 2
            SQRSIN = 3
            MAXWELL = 4

        class LfAugType(_NamedEnum):
            Z = 0
            X = 1

        class HlambdaModeType(_NamedEnum):
            CONST = 0
            DYNAMICAL = 1

        class InitGuessHf(_NamedEnum):
            EXP = 0
            COS = 1
            COS_SET = 2
            SIN = 3
            SIN_SET = 4

        class FType(_NamedEnum):
            SM = 0
            RE = 1
            SS = 2

        class PropagationConfiguration(ConfigurationBase):
            def __init__(self):
                super().__init__()
                self._define(m=0.5, nch=64, nt=420000, E0=71.54, t0=300e-15, sigma=50e-15, nu_L=0.29297e15)

        def __init__(self):
            super().__init__()
            F = TaskRootConfiguration.FitterConfiguration
            self._define(F_type=F.FType.SM, w_list=[], w_min=-2.0, w_max=2.0,
                         propagation=F.PropagationConfiguration(), k_E=1e29, lamb=4e14, pow=0.8, epsilon=1e-15,
                         impulses_number=2, delay=600e-15, mod_log=500, iter_max=-1, iter_mid_1=0, iter_mid_2=1,
                         q=0.0, h_lambda=0.0066, h_lambda_mode=F.HlambdaModeType.CONST, pcos=1.0, hf_hide=True,
                         Em=1.5)

    def __init__(self):
        super().__init__()
        T = TaskRootConfiguration
        F = T.FitterConfiguration
        self._define(task_type=T.TaskType.TRANS_WO_CONTROL, pot_type=T.PotentialType.MORSE,
                     wf_type=T.WaveFuncType.MORSE, hamil_type=T.HamilType.NTRIV, lf_aug_type=F.LfAugType.Z,
                     init_guess=F.InitGuess.ZERO, init_guess_hf=F.InitGuessHf.EXP, fitter=F(), run_id="no_id",
                     nb=1, nlevs=2, T=600e-15, L=5.0, np=1024, De=20000.0, De_e=10000.0, Du=20000.0, x0p=-0.17,
                     x0=0.0, p0=0.0, a=1.0, a_e=1.0, U=0.0, W=0.0, delta=1.0, t0_auto=False, nt_auto=False,
                     sigma_auto=False, nu_L_auto=False)


class ReportRootConfiguration(ConfigurationBase):
    """Report file of ``--json_rep`` (config.py:313-426 of the reference): one tree per output family, selected
    by the key prefix ("table." / "plot."); un-prefixed keys (out_path, plotting_flag) are shared."""

    class ReportFitterConfiguration(ConfigurationBase):
        class OutputType(_NamedEnum):
            ALL = 0
            TABLES = 1
            TABLES_ITER = 2
            PLOTS = 3
            NONE = 4

        class ReportPropagationConfiguration(ConfigurationBase):
            pass

        class ReportTablePropagationConfiguration(ReportPropagationConfiguration):
            def __init__(self, suffix=None):
                super().__init__("table.")
                self._define(tab_abs="tab_abs_{level}.csv", tab_real="tab_real_{level}.csv",
                             tab_tvals="tab_tvals_{level}.csv", tab_tvals_fit="tab_tvals_fit.csv", lmin=0,
                             mod_fileout=100)

        class ReportPlotPropagationConfiguration(ReportPropagationConfiguration):
            def __init__(self, suffix=None):
                super().__init__("plot.")
                self._define(lmin=0, mod_plotout=100, mod_update=20, number_plotout=15)

        def __init__(self, key_prefix=""):
            super().__init__(key_prefix)
            R = ReportRootConfiguration.ReportFitterConfiguration
            self._define(propagation=R.ReportPropagationConfiguration(key_prefix), out_path="output",
                         table_glob_path="", plotting_flag=R.OutputType.ALL)

    class ReportTableFitterConfiguration(ReportFitterConfiguration):
        def __init__(self, suffix=None):
            super().__init__("table.")
            self._define(propagation=ReportRootConfiguration.ReportFitterConfiguration.ReportTablePropagationConfiguration(),
                         tab_iter="tab_iter.csv", tab_iter_E="tab_iter_E.csv", imin=0, imod_fileout=1)

    class ReportPlotFitterConfiguration(ReportFitterConfiguration):
        def __init__(self, suffix=None):
            super().__init__("plot.")
            self._define(propagation=ReportRootConfiguration.ReportFitterConfiguration.ReportPlotPropagationConfiguration(),
                         imin=0, imod_plotout=1, inumber_plotout=15)

    def __init__(self, key_prefix=""):
        super().__init__(key_prefix)
        self._define(fitter=ReportRootConfiguration.ReportFitterConfiguration(key_prefix))


class ReportTableRootConfiguration(ReportRootConfiguration):
    def __init__(self):
        super().__init__("table.")
        self._define(fitter=ReportRootConfiguration.ReportTableFitterConfiguration())


class ReportPlotRootConfiguration(ReportRootConfiguration):
    def __init__(self):
        super().__init__("plot.")
        self._define(fitter=ReportRootConfiguration.ReportPlotFitterConfiguration())
