"""Test-only stand-in for the `trimesh` package (absent from the image), with just what the reference's
tests/test_mesh.py:12-29 uses: Trimesh(vertices, faces), util.concatenate, .process() (merge duplicate
vertices) and .is_watertight (every edge belongs to exactly two faces, trimesh's definition)."""
import numpy as np


class Trimesh:
    def __init__(self, vertices, faces):
        self.vertices = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
        self.faces = np.asarray(faces, dtype=np.int64).reshape(-1, 3)

    def process(self):
        # trimesh merges vertices that agree to 1e-8 of the mesh scale
        scale = float(np.abs(self.vertices).max()) if len(self.vertices) else 1.0
        key = np.round(self.vertices / (max(scale, 1e-300) * 1e-8)).astype(np.int64)
        _, first, inverse = np.unique(key, axis=0, return_index=True, return_inverse=True)
        self.vertices = self.vertices[first]
        self.faces = inverse.reshape(-1)[self.faces]
        degenerate = (self.faces[:, 0] == self.faces[:, 1]) | (self.faces[:, 1] == self.faces[:, 2]) | \
                     (self.faces[:, 0] == self.faces[:, 2])
        self.faces = self.faces[~degenerate]
        return self

    @property
    def is_watertight(self):
        if len(self.faces) == 0:
            return False
        f = self.faces
        edges = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        edges.sort(axis=1)
        _, counts = np.unique(edges, axis=0, return_counts=True)
        return bool(np.all(counts == 2))


class util:  # noqa: N801 - mirrors trimesh.util
    @staticmethod
    def concatenate(a, b):
        return Trimesh(np.concatenate([a.vertices, b.vertices]),
                       np.concatenate([a.faces, b.faces + len(a.vertices)]))
