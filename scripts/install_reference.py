#!/usr/bin/env python
"""Place an UNMODIFIED copy of the reference's hot-path packages (tgcn/, gcn/ -- Python sources only) under
baseline/_ref/ so that `bench.py --impl reference` can time the reference's own layer classes (cpu_baseline.kind
"reference") on a box where /root/reference does not exist.  baseline/_ref/ is git-ignored (never part of this
repository's history) but travels with the gpurun snapshot.  The reference has no setup.py / pyproject, so there is
nothing for pip to install: the two package directories are copied verbatim."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def install(src="/root/reference", dst=os.path.join(ROOT, "baseline", "_ref")):
    if not os.path.isfile(os.path.join(src, "tgcn", "nn", "gcn.py")):
        return None
    for pkg in ("tgcn", "gcn"):
        out = os.path.join(dst, pkg)
        if os.path.isdir(out):
            shutil.rmtree(out)
        shutil.copytree(os.path.join(src, pkg), out, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.mat", "*.npy"))
    return dst


if __name__ == "__main__":
    print(install(*sys.argv[1:]))
