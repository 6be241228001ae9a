"""Accuracy of the canonical transcendental functions (oracle/cc_math_ref.h, mirrored bit
for bit by codecad_b200/csrc/cc_math.cuh) against float64 libm.  Device-free."""
import numpy as np

import oracle


def _ulp(x):
    return np.spacing(np.abs(x).astype(np.float32)).astype(np.float64)


def test_atan2_accuracy():
    rng = np.random.default_rng(1)
    y = rng.uniform(-100, 100, 200000).astype(np.float32)
    x = rng.uniform(-100, 100, 200000).astype(np.float32)
    got, _ = oracle.math_probe(0, y, x)
    want = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    assert np.max(np.abs(got - want)) < 4e-7           # ~1.5 ulp at pi
    # axes and the origin
    got, _ = oracle.math_probe(0, np.array([0, 0, 1, -1, 0], np.float32), np.array([1, -1, 0, 0, 0], np.float32))
    assert np.allclose(got, [0, np.pi, np.pi / 2, -np.pi / 2, 0], atol=1e-7)


def test_sincos_accuracy():
    x = np.linspace(-200, 200, 400001).astype(np.float32)
    s, c = oracle.math_probe(1, x)
    assert np.max(np.abs(s - np.sin(x.astype(np.float64)))) < 2.5e-7
    assert np.max(np.abs(c - np.cos(x.astype(np.float64)))) < 2.5e-7
    assert np.max(np.abs(s.astype(np.float64) ** 2 + c.astype(np.float64) ** 2 - 1)) < 5e-7


def test_acos_accuracy():
    x = np.linspace(-1, 1, 200001).astype(np.float32)
    got, _ = oracle.math_probe(2, x)
    assert np.max(np.abs(got - np.arccos(x.astype(np.float64)))) < 6e-7


def test_fmod_and_remainder():
    rng = np.random.default_rng(2)
    x = rng.uniform(0, 50, 100000).astype(np.float32)
    y = rng.uniform(0.05, 7, 100000).astype(np.float32)
    got, _ = oracle.math_probe(3, x, y)
    assert np.all((got >= 0) & (got < y))
    want = np.fmod(x.astype(np.float64), y.astype(np.float64))
    err = np.abs(got - want)
    assert np.all(np.minimum(err, np.abs(err - y)) < 1e-5)      # equal modulo y up to rounding
    xs = rng.uniform(-50, 50, 100000).astype(np.float32)
    got, _ = oracle.math_probe(4, xs, y)
    assert np.all(np.abs(got) <= y * 0.5000001)
    want = np.remainder(xs.astype(np.float64) + y / 2.0, y.astype(np.float64)) - y / 2.0
    err = np.abs(got - want)
    assert np.all(np.minimum(err, np.abs(err - y)) < 1e-5)
    # remainder(x, inf) == x  (unsafe.cl:1-6 relies on it)
    got, _ = oracle.math_probe(4, xs[:100], np.full(100, np.inf, np.float32))
    assert np.array_equal(got, xs[:100])
