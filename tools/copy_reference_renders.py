#!/usr/bin/env python3
"""Provenance of tests/golden/reference_renders/: the 32 baseline pictures of the reference's test
suite (/root/reference/tests/baseline/rendered_<shape>.png, compared by its tests/test_image.py:16-28),
copied verbatim as golden DATA — the only stored outputs of the reference's OpenCL evaluate().
No reference source code is copied.  Run in the build container (the GPU box has no /root/reference):

    python tools/copy_reference_renders.py
"""
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/tests/baseline"
DST = os.path.join(ROOT, "tests", "golden", "reference_renders")

if __name__ == "__main__":
    os.makedirs(DST, exist_ok=True)
    names = sorted(n for n in os.listdir(SRC) if n.startswith("rendered_") and n.endswith(".png"))
    for n in names:
        shutil.copyfile(os.path.join(SRC, n), os.path.join(DST, n))
    print("copied %d pictures" % len(names))
