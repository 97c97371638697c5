"""The C-ABI shared library: it loads, exports every symbol include/qcontrol_b200.h declares, and
fails loudly (no CPU fallback) where a GPU is required.  No compute calls here -- CPU only."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "qcontrol_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qc_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from qcontrol_b200 import _lib
    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from qcontrol_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (qc_[a-z0-9_]+)\b", out))
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, missing
    lib = C.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        getattr(lib, s)
    assert lib.qc_abi_version() == _lib.ABI_VERSION


def test_library_is_built_for_sm_100a_with_lineinfo():
    from qcontrol_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_argument_errors_are_reported_not_swallowed():
    from qcontrol_b200 import _lib
    lib = _lib.lib
    h = C.c_void_p()
    assert lib.qc_plan_create(None, None, C.byref(h)) == -1
    assert b"NULL" in lib.qc_last_error()
    assert lib.qc_sweep(None, None) == -1
    assert lib.qc_set_state(None, None) == -1


def test_no_cpu_fallback_without_a_gpu():
    """On a box without a CUDA device the product path raises; it never computes on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from qcontrol_b200 import _lib, engine
    with pytest.raises(_lib.QcError) as e:
        engine.Context(0)
    assert e.value.code == -2          # QC_ERR_CUDA


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "qcontrol_b200")
    for base, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "qc_oracle" not in text, f
