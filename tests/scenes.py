"""Loader for tests/golden/scenes.npz (made by tools/make_fixtures.py from the
reference's own node compiler)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Scene:
    def __init__(self, name, words, meta):
        self.name = name
        self.words = np.ascontiguousarray(words, dtype=np.float32)
        self.dimension = int(meta[0])
        self.box_a = tuple(float(v) for v in meta[1:4])
        self.box_b = tuple(float(v) for v in meta[4:7])
        self.feature_size = float(meta[7])

    def compiled(self):
        from codecad_b200 import CompiledScene
        return CompiledScene(self.words, self.dimension, self.box_a, self.box_b, self.feature_size, self.name)

    def grid(self, n, margin=1.02):
        """cubic n^3 grid covering the bounding box: SURVEY.md 8(d) convention."""
        size = max(b - a for a, b in zip(self.box_a, self.box_b))
        step = margin * size / n
        mid = [(a + b) / 2 for a, b in zip(self.box_a, self.box_b)]
        corner = [m - step * (n - 1) / 2 for m in mid]
        return np.array(corner, np.float64).astype(np.float32), np.float32(step)


FILES = ("scenes.npz", "forest_scenes.npz", "column_scenes.npz")  # tools/make_fixtures.py [--forests | --columns]


def load_scenes():
    out = {}
    for f in FILES:
        z = np.load(os.path.join(GOLDEN, f))
        for n in sorted(k[:-6] for k in z.files if k.endswith(".words")):
            out[n] = Scene(n, z[n + ".words"], z[n + ".meta"])
    return out


ALL_NAMES = sorted(load_scenes())
COLUMN_NAMES = [n for n in ALL_NAMES if n.startswith(("col_", "colr_"))]
FOREST_NAMES = [n for n in ALL_NAMES if n.startswith("forest_") or n.startswith("cfg_synthetic")]
