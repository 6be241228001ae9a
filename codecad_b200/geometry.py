"""Minimal Vector / BoundingBox with the reference's semantics
(/root/reference/codecad/util/geometry.py:8-117,119-200): float64 namedtuples.  They
compare and hash like the reference's (plain tuples), so results can be mixed freely with
`codecad.util.Vector` when the reference package is present."""
import collections
import math

import numpy as np

FLOAT4 = np.dtype([(n, np.float32) for n in "xyzw"])   # pyopencl.cltypes.float4
FLOAT2 = np.dtype([(n, np.float32) for n in "xy"])
UCHAR4 = np.dtype([(n, np.uint8) for n in "xyzw"])


class Vector(collections.namedtuple("Vector", "x y z")):
    __slots__ = ()

    def __new__(cls, x, y, z=0):
        return super().__new__(cls, x, y, z)

    @classmethod
    def splat(cls, value):
        return cls(value, value, value)

    def __add__(self, other):
        return Vector(self.x + other.x, self.y + other.y, self.z + other.z)

    def __sub__(self, other):
        return Vector(self.x - other.x, self.y - other.y, self.z - other.z)

    def __mul__(self, other):
        return Vector(self.x * other, self.y * other, self.z * other)

    def __truediv__(self, other):
        return Vector(self.x / other, self.y / other, self.z / other)

    def __neg__(self):
        return Vector(-self.x, -self.y, -self.z)

    def __abs__(self):
        return math.sqrt(self.x * self.x + self.y * self.y + self.z * self.z)

    def abs_squared(self):
        return self.dot(self)

    def dot(self, other):
        return self.x * other.x + self.y * other.y + self.z * other.z

    def cross(self, other):
        return Vector(self.y * other.z - self.z * other.y, self.z * other.x - self.x * other.z,
                      self.x * other.y - self.y * other.x)

    def normalized(self):
        return self / abs(self)

    def elementwise_mul(self, other):
        return Vector(self.x * other.x, self.y * other.y, self.z * other.z)

    def elementwise_div(self, other):
        return Vector(self.x / other.x, self.y / other.y, self.z / other.z)

    def max(self, other=None):
        if other is None:
            return max(self.x, self.y, self.z)
        return Vector(max(self.x, other.x), max(self.y, other.y), max(self.z, other.z))

    def min(self, other=None):
        if other is None:
            return min(self.x, self.y, self.z)
        return Vector(min(self.x, other.x), min(self.y, other.y), min(self.z, other.z))

    def applyfunc(self, f):
        return Vector(f(self.x), f(self.y), f(self.z))

    def flattened(self):
        return Vector(self.x, self.y, 0)

    def as_float2(self):
        return np.array((self.x, self.y), dtype=FLOAT2)

    def as_float4(self, w=0):
        """numpy structured float4 scalar, rounded to fp32 (geometry.py:98-99)."""
        return np.array((self.x, self.y, self.z, w), dtype=FLOAT4)


class BoundingBox(collections.namedtuple("BoundingBox", "a b")):
    __slots__ = ()

    def size(self):
        return self.b - self.a

    def midpoint(self):
        return (self.a + self.b) / 2

    def expanded_additive(self, expansion):
        e = Vector.splat(expansion)
        return BoundingBox(self.a - e, self.b + e)

    def flattened(self):
        return BoundingBox(self.a.flattened(), self.b.flattened())


def as_vector(v):
    if isinstance(v, Vector):
        return v
    return Vector(*(tuple(v)[:3]))
