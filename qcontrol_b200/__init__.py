"""qcontrol_b200 -- B200-native propagation hot path of irenemizus/qcontrol.

Importing the package loads the CUDA library (qcontrol_b200/lib/libqcontrol_b200.so);
there is no CPU fallback: a missing library is an ImportError, a missing GPU is a
QcError from the first context creation.
"""
from . import _lib            # noqa: F401  (fails loudly when the CUDA library is absent)
from .engine import Context, Plan  # noqa: F401

__version__ = "0.1.0"
