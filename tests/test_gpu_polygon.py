"""GPU 2-D outline path: cc_process_polygon / cc_polygon_blocks and the
`rendering.polygon2d.polygon` mirror against the CPU oracle — bit-exact vertices, identical links,
identical polygons."""
import ctypes

import numpy as np
import pytest

import oracle
from oracle import host
from scenes import ALL_NAMES
from test_polygon_oracle import canonical, signed_area

pytestmark = pytest.mark.gpu

NAMES_2D = [n for n in ALL_NAMES if n.startswith("dsdf2d_")]


@pytest.fixture(scope="module")
def cb():
    import codecad_b200
    from codecad_b200 import _lib
    _lib.init(0)
    old = _lib.check(_lib.lib().cc_set_jit_mode(0))
    yield codecad_b200
    _lib.check(_lib.lib().cc_set_jit_mode(old))


@pytest.mark.parametrize("name", NAMES_2D)
def test_process_polygon_kernel_bit_exact(cb, scenes, name):
    """The kernel with the reference's per-launch semantics on device buffers."""
    from codecad_b200 import _lib
    from codecad_b200.cl_util.buffer import ProgramBuffer
    L = _lib.lib()
    s = scenes[name]
    gx, gy = 37, 29                                        # ragged
    corner, step = s.grid(40)
    field = oracle.grid_eval(s.words, corner, step, (gx, gy, 1))
    want_v, want_l, want_s = oracle.process_polygon(corner[:2], step, field[:, :, 0, :])
    cells = 2 * (gx - 1) * (gy - 1)
    prog = ProgramBuffer(s.words)
    bufs = [ctypes.c_void_p() for _ in range(5)]
    sizes = [gx * gy * 16, cells * 8, cells * 4, (gx + gy - 2) * 4, 4]
    for b, n in zip(bufs, sizes):
        _lib.check(L.cc_buffer_alloc(n, ctypes.byref(b)))
    try:
        d_field, d_v, d_l, d_s, d_c = bufs
        c3 = (ctypes.c_float * 3)(*np.asarray(corner, np.float32)[:3])
        _lib.check(L.cc_grid_eval(prog.handle, c3, ctypes.c_float(step), gx, gy, 1, 0, _lib.LAYOUT_INDEX3_FLOAT4, d_field, None))
        _lib.check(L.cc_memset_async(d_v, 0, sizes[1], None))
        _lib.check(L.cc_memset_async(d_c, 0, 4, None))
        c2 = (ctypes.c_float * 2)(*np.asarray(corner, np.float32)[:2])
        _lib.check(L.cc_process_polygon(c2, ctypes.c_float(step), gx - 1, gy - 1, d_field, d_v, d_l, d_s, gx + gy - 2, d_c, None))
        got_v, got_l = np.empty((cells, 2), np.float32), np.empty(cells, np.uint32)
        got_s, got_c = np.empty(gx + gy - 2, np.uint32), np.empty(1, np.uint32)
        for h, d in ((got_v, d_v), (got_l, d_l), (got_s, d_s), (got_c, d_c)):
            _lib.check(L.cc_memcpy_d2h_async(h.ctypes.data, d, h.nbytes, None))
        _lib.check(L.cc_synchronize())
    finally:
        for b in bufs:
            L.cc_buffer_free(b)
    assert np.array_equal(got_l, want_l)
    assert np.array_equal(got_v, want_v)
    assert int(got_c[0]) == len(want_s) and np.array_equal(got_s[:len(want_s)], want_s)


@pytest.mark.parametrize("grid", [None, 24, 9, 5])
@pytest.mark.parametrize("name", NAMES_2D)
def test_polygon_matches_oracle(cb, scenes, name, grid):
    from codecad_b200.rendering import polygon2d
    s = scenes[name]
    got = list(polygon2d.polygon(s.compiled(), grid))
    want = host.polygon(s.words, s.box_a, s.box_b, s.feature_size, 128 if grid is None else grid)
    assert len(got) == len(want) >= 1
    if grid is None:
        # one box: the reference's order (increasing first triangle) and starting vertices
        assert [[tuple(v) for v in c] for c in got] == [[tuple(v) for v in c] for c in want]
    else:
        assert canonical(got) == canonical(want)


def test_fine_outline_of_the_gear(cb, scenes):
    """A finer resolution than the default: ~10 k vertices over hundreds of boxes."""
    from codecad_b200 import CompiledScene
    from codecad_b200.rendering import polygon2d
    s = scenes["dsdf2d_gear"]
    fine = CompiledScene(s.words, 2, s.box_a, s.box_b, s.feature_size / 16, s.name)
    got = list(polygon2d.polygon(fine, 32))
    want = host.polygon(s.words, s.box_a, s.box_b, s.feature_size / 16, 32)
    assert canonical(got) == canonical(want)
    assert len(got) == 1 and len(got[0]) > 5000
    assert signed_area(got[0]) == pytest.approx(78.0, rel=0.02)
