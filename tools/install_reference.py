#!/usr/bin/env python3
"""Put the reference's *Python half* where the GPU box can see it: baseline/_ref/ (git-ignored,
but it travels with the gpurun snapshot like the built .so files).

    baseline/_ref/codecad/           the reference package, unmodified
    baseline/_ref/reference_tests/   its test-suite (tests/*.py, *.cl, baseline PNGs), unmodified

`pip install --target baseline/_ref /root/reference` cannot work offline: setup.cfg has
`setup_requires=pytest-runner`, which setuptools tries to download (recorded in DESIGN.md §8).  The
package is pure Python plus .cl data files, so the install is a copy of the tree.  Nothing here is
committed: the copy exists so that tests/test_reference_suite.py can run the reference's OWN tests
against the CUDA path on the B200 (`codecad_b200.dropin.load()` swaps the OpenCL layer).
Runs only where /root/reference exists (this container)."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("CODECAD_REFERENCE", "/root/reference")
DEST = os.path.join(REPO, "baseline", "_ref")


def install(force=False):
    if not os.path.isdir(os.path.join(REF, "codecad")):
        return None
    stamp = os.path.join(DEST, ".installed_from")
    if not force and os.path.exists(stamp) and open(stamp).read().strip() == REF:
        return DEST
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc")
    for src, dst in (("codecad", "codecad"), ("tests", "reference_tests"), ("examples", "examples")):
        target = os.path.join(DEST, dst)
        if os.path.isdir(target):
            shutil.rmtree(target)
        shutil.copytree(os.path.join(REF, src), target, ignore=ignore)
    with open(stamp, "w") as f:
        f.write(REF + "\n")
    return DEST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
