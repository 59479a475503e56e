"""ORACLE / test infrastructure: vendor the UNMODIFIED reference package to ``oracle/_ref`` so that it travels to the GPU box.

    python -m oracle.vendor_ref            (run in the build container; __graft_entry__.build() calls it)

The reference is a pure-Python poetry project without a build step (``/root/reference/pyproject.toml``); "installing" it is
copying its ``acquisition_focus`` package.  The copy goes to ``oracle/_ref/acquisition_focus`` - listed in ``.gitignore`` (no
reference source ever enters the history) but not in ``.gpurunignore``, so the GPU box receives it like a built ``.so``.  It is
used ONLY by ``bench.py``'s CPU arm (``--impl reference`` / ``cpu_baseline``, ``kind: "reference"``) to time the reference's own
code on the box's host cores; ``/root/reference`` itself does not exist there.  Only the modules on this path are vendored
(utils / models / functional / running; no datasets, notebooks, artifacts).
"""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(os.environ.get("AFB_REFERENCE_SRC", "/root/reference"), "acquisition_focus")
DST = os.path.join(ROOT, "oracle", "_ref", "acquisition_focus")
KEEP = ("__init__.py", "utils", "models", "functional", "running")


def vendor(force: bool = False) -> str | None:
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None          # GPU box / no reference mounted: use what travelled
    if os.path.isdir(DST):
        shutil.rmtree(DST)           # always a fresh copy of what /root/reference holds now
    os.makedirs(DST, exist_ok=True)
    for name in KEEP:
        s, d = os.path.join(SRC, name), os.path.join(DST, name)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.ipynb", "*.pth"), dirs_exist_ok=True)
        elif os.path.exists(s):
            shutil.copy2(s, d)
    with open(os.path.join(ROOT, "oracle", "_ref", "README"), "w") as fh:
        fh.write("Unmodified copy of /root/reference/acquisition_focus made by oracle/vendor_ref.py; git-ignored; used only by the CPU arm of bench.py.\n")
    return DST


if __name__ == "__main__":
    print(vendor(force="--force" in sys.argv))
