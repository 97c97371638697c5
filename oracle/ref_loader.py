"""Load the UNMODIFIED reference modules (irenemizus/qcontrol) from a directory  --  TEST INFRASTRUCTURE ONLY.

Two roots are used: ``/root/reference`` in the build container (fixture generation, tests/golden/gen_golden.py) and
``oracle/_ref/`` -- a git-ignored verbatim copy made by ``oracle/build_ref.py`` so that the reference itself can be
timed on the GPU box's host cores (``bench.py --impl reference``, ``cpu_baseline.kind = "reference"``).  Nothing of
the reference is stored in the repository's history and the product package never imports this module.

One load-time shim, in memory: ``ndarray.itemset((i, j), v)`` (removed in NumPy 2) is rewritten to index assignment
in ``hamil_2d.py`` / ``task_manager.py`` before ``exec`` (13 call sites, hamil_2d.py:111-178, task_manager.py:762).
"""
from __future__ import annotations

import os
import re
import sys
import types

MODULES = ("math_base", "phys_base", "psi_basis", "hamil_2d", "propagation", "task_manager", "fitter", "config",
           "reporter", "tools", "grid_setup")
DEFAULT_ROOT = "/root/reference"
SHIPPED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available_root():
    """The build container's mount if present, else the shipped copy, else None."""
    for root in (DEFAULT_ROOT, SHIPPED_ROOT):
        if all(os.path.exists(os.path.join(root, m + ".py")) for m in MODULES):
            return root
    return None


def load_reference(root=DEFAULT_ROOT):
    if root not in sys.path:
        sys.path.insert(0, root)
    shim = re.compile(r"(\w+)\.itemset\(\((.+?),\s*(.+?)\),\s*(.+)\)\s*(#.*)?$", re.M)

    def load_shimmed(name):
        src = open(os.path.join(root, name + ".py")).read()
        src = shim.sub(lambda m: "%s[%s, %s] = %s" % (m.group(1), m.group(2), m.group(3), m.group(4)), src)
        assert "itemset" not in src, name
        mod = types.ModuleType(name)
        mod.__file__ = os.path.join(root, name + ".py")
        sys.modules[name] = mod
        exec(compile(src, mod.__file__, "exec"), mod.__dict__)
        return mod

    import math_base, phys_base, psi_basis  # noqa: F401
    load_shimmed("hamil_2d")
    import propagation  # noqa: F401
    load_shimmed("task_manager")
    import fitter, config, reporter  # noqa: F401
    return types.SimpleNamespace(root=root, **{k: sys.modules[k] for k in
                                               ("math_base", "phys_base", "psi_basis", "hamil_2d", "propagation",
                                                "task_manager", "fitter", "config", "reporter")})
