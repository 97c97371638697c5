"""Recipe for ``oracle/_ref/``: a verbatim, git-ignored copy of the reference modules on the hot path.

    python oracle/build_ref.py            (run by __graft_entry__.build() when /root/reference is present)

The reference is pure Python (no build system, no setup.py), so "building" it is copying the eleven modules the
propagation path imports, byte for byte, to a directory that travels to the GPU box with the snapshot (listed in
.gitignore, not in .gpurunignore).  ``oracle/ref_loader.py`` imports them from there; ``bench.py --impl reference``
then times the reference's own ``PropagationSolver.step`` -- instrumentation and per-step Newton table included, as the
reference always computes them -- on the box's host cores.  A SHA-256 manifest is written beside the copy so that a
stale or edited copy is detected (``verify``)."""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.ref_loader import DEFAULT_ROOT, MODULES, SHIPPED_ROOT  # noqa: E402


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(src=DEFAULT_ROOT, dst=SHIPPED_ROOT):
    if not os.path.isdir(src):
        return False
    os.makedirs(dst, exist_ok=True)
    manifest = {}
    for m in MODULES:
        shutil.copyfile(os.path.join(src, m + ".py"), os.path.join(dst, m + ".py"))
        manifest[m + ".py"] = _sha(os.path.join(dst, m + ".py"))
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": "irenemizus/qcontrol (unmodified copies of %s)" % src, "sha256": manifest}, f, indent=1)
    return True


def verify(dst=SHIPPED_ROOT):
    """True when every module of the copy matches its manifest entry."""
    try:
        with open(os.path.join(dst, "MANIFEST.json")) as f:
            manifest = json.load(f)["sha256"]
        return all(_sha(os.path.join(dst, name)) == h for name, h in manifest.items()) and len(manifest) == len(MODULES)
    except (OSError, KeyError, ValueError):
        return False


if __name__ == "__main__":
    ok = build()
    print("oracle/_ref: %s" % ("copied %d modules" % len(MODULES) if ok else "no reference mount, nothing done"))
