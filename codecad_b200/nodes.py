"""`codecad.nodes` entry points used by the hot path (nodes/program.py:74-84).

The node compiler itself (shape tree -> scheduled float32 instruction stream) is the
reference's own Python and stays there; what changes is where the words go:
`make_program_buffer()` hands them to libcodecad_b200, which decodes and uploads them.

A shape can also be a `CompiledScene`: program words produced earlier by the reference
compiler (tests/golden/scenes.npz, or a cache kept by the caller — the reference's
scheduler costs 0.1-38 s per scene) together with the two facts the
drivers ask a shape for, its dimension and bounding box.
"""
import numpy as np

from .cl_util.buffer import ProgramBuffer
from .geometry import BoundingBox, Vector


class CompiledScene:
    """Pre-compiled scene: words + dimension + bounding box (+ feature size)."""

    def __init__(self, words, dimension, box_a, box_b, feature_size=None, name=None):
        self.words = np.ascontiguousarray(words, dtype=np.float32)
        self._dimension = int(dimension)
        self._box = BoundingBox(Vector(*box_a), Vector(*box_b))
        self._feature_size = feature_size
        self.name = name
        self._buffer = None

    def dimension(self):
        return self._dimension

    def bounding_box(self):
        return self._box

    def feature_size(self):
        return self._feature_size

    def program_buffer(self):
        if self._buffer is None:
            self._buffer = ProgramBuffer(self.words)
        return self._buffer


def make_program(shape):
    """float32 instruction words of `shape` (nodes/program.py:74-76)."""
    if isinstance(shape, CompiledScene):
        return shape.words
    if isinstance(shape, np.ndarray):
        return np.ascontiguousarray(shape, dtype=np.float32)
    try:
        from codecad.nodes import program as _ref_program  # the reference's compiler
    except ImportError as exc:
        raise TypeError(
            "make_program() needs a CompiledScene / word array, or the reference `codecad` "
            "package on sys.path to compile %r" % (shape,)) from exc
    return _ref_program.make_program(shape)


def make_program_buffer(shape):
    """Device-resident program for `shape` (nodes/program.py:79-84)."""
    if isinstance(shape, ProgramBuffer):
        return shape
    if isinstance(shape, CompiledScene):
        return shape.program_buffer()
    return ProgramBuffer(make_program(shape))
