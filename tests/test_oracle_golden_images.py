"""The oracle's evaluate() against the reference's own stored outputs.

tests/golden/reference_renders/*.png are the 32 baseline images of the reference's test suite
(tests/baseline/rendered_<shape>.png, compared by tests/test_image.py:16-28 with
tests/tools.py:64-79: mean squared error of the [0,1] RGB arrays <= 1e-3).  They were rendered by the
reference's OpenCL evaluate() through its ray caster (3-D shapes) and bitmap kernel (2-D shapes):
the only real outputs of the reference's device code that exist.  oracle/render.py + the renderer
restatement in oracle/sdf_oracle.c reproduce every one of them from the fixture program words,
which pins the oracle's interpreter and all op restatements these 32 shapes use."""
import os

import numpy as np
import pytest
from PIL import Image

from oracle import render
from scenes import ALL_NAMES

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_renders")
NAMES = [n for n in ALL_NAMES if n.startswith("dsdf2d_") or n.startswith("dsdf3d_")]


def _mse(a, b):
    d = a.astype(np.float32) / 255 - b.astype(np.float32) / 255
    return float(np.mean(d * d))


def test_all_32_reference_shapes_are_covered():
    assert len(NAMES) == 32
    assert sorted("rendered_%s.png" % n.split("_", 1)[1] for n in NAMES) == sorted(os.listdir(GOLD))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference_image(scenes, name):
    s = scenes[name]
    gold = np.asarray(Image.open(os.path.join(GOLD, "rendered_%s.png" % name.split("_", 1)[1])).convert("RGB"))
    img = render.render(s.words, s.dimension, s.box_a, s.box_b, (gold.shape[1], gold.shape[0]))
    assert img.shape == gold.shape
    mse = _mse(img, gold)
    assert mse <= 1e-3, mse                       # the reference's own tolerance (tests/tools.py:74)
    # in fact the images are nearly identical: at most a few hundred of the 786 432 pixels differ
    assert mse <= 5e-4
    assert (np.abs(img.astype(int) - gold.astype(int)).max(axis=-1) > 8).mean() < 0.003
