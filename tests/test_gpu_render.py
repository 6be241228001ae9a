"""GPU image renderers: cc_ray_caster / cc_bitmap through the
`rendering.image` mirror, against (a) the reference's own 32 golden PNGs (tests/baseline of the
reference, its tolerance: tests/tools.py:64-79) and (b) the CPU oracle, bit for bit."""
import os

import numpy as np
import pytest
from PIL import Image

from oracle import render as oracle_render
from scenes import ALL_NAMES

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_renders")
NAMES = [n for n in ALL_NAMES if n.startswith("dsdf2d_") or n.startswith("dsdf3d_")]


@pytest.fixture(scope="module")
def cb():
    import codecad_b200
    from codecad_b200 import _lib
    _lib.init(0)
    old = _lib.check(_lib.lib().cc_set_jit_mode(0))
    yield codecad_b200
    _lib.check(_lib.lib().cc_set_jit_mode(old))


def _mse(a, b):
    d = a.astype(np.float32) / 255 - b.astype(np.float32) / 255
    return float(np.mean(d * d))


@pytest.mark.parametrize("name", NAMES)
def test_reference_golden_image(cb, scenes, name):
    """The reference's tests/test_image.py:16-28, with the CUDA path rendering."""
    from codecad_b200.rendering import image
    s = scenes[name]
    gold = np.asarray(Image.open(os.path.join(GOLD, "rendered_%s.png" % name.split("_", 1)[1])).convert("RGB"))
    size = (gold.shape[1], gold.shape[0])
    got = image.render_pixels(s.compiled(), size)
    assert got.shape == gold.shape and got.dtype == np.uint8
    assert _mse(got, gold) <= 1e-3                      # the reference's own check
    want = oracle_render.render(s.words, s.dimension, s.box_a, s.box_b, size)
    assert np.array_equal(got, want)                    # and bit-exact against the oracle


@pytest.mark.parametrize("options", [1, 2, 3])
@pytest.mark.parametrize("name", ["dsdf3d_csg_thing", "dsdf3d_torus", "cfg_planetary"])
def test_ray_caster_render_options(cb, scenes, name, options):
    from codecad_b200.rendering import ray_caster
    s = scenes[name]
    scene = s.compiled()
    size = (203, 151)                                   # ragged: partial 8x4 warp tiles
    cam = ray_caster.get_camera_params(scene.bounding_box(), size, None)
    stats = {}
    got = ray_caster.render(scene, size=size, options=ray_caster.RenderOptions(options), stats=stats, *cam)
    want = oracle_render.ray_cast(s.words, s.box_a, s.box_b, size, options=options)
    assert np.array_equal(got, want)
    assert stats["evaluations"] >= size[0] * size[1] and stats["ms"] > 0


@pytest.mark.parametrize("size", [(1, 1), (7, 3), (64, 48)])
def test_small_images(cb, scenes, size):
    from codecad_b200.rendering import image
    for name in ("dsdf3d_box", "dsdf2d_gear"):
        s = scenes[name]
        got = image.render_pixels(s.compiled(), size)
        assert np.array_equal(got, oracle_render.render(s.words, s.dimension, s.box_a, s.box_b, size))


def test_view_angle_and_large_program(cb, scenes):
    """cfg_synthetic500: 12 k words, program staged through shared memory or the constant bank."""
    from codecad_b200.rendering import image
    s = scenes["cfg_synthetic500"]
    size = (96, 64)
    got = image.render_pixels(s.compiled(), size, view_angle=30)
    want = oracle_render.ray_cast(s.words, s.box_a, s.box_b, size, view_angle=30)
    assert np.array_equal(got, want)


# ---- scene-specialised renderers (NVRTC): same render body, same op library -> same bytes -------

@pytest.mark.parametrize("name", NAMES + ["cfg_planetary", "cfg_menger_sponge", "col_assembly", "dsdf3d_mirror_3d"])
def test_specialised_renderers_bit_exact(cb, scenes, name):
    from codecad_b200 import CompiledScene
    from codecad_b200.cl_util.buffer import ProgramBuffer
    from codecad_b200.rendering import image
    s = scenes[name]
    scene = s.compiled()
    prog = scene.program_buffer()
    sink = ProgramBuffer.SINK_BITMAP if s.dimension == 2 else ProgramBuffer.SINK_RAY
    assert prog.specialize(1, sink) > 0 and prog.use_specialized(True)
    size = (333, 251)
    from codecad_b200 import _lib
    launches0, _ = _lib.counters()
    got = image.render_pixels(scene, size)
    assert _lib.counters()[0] == launches0 + 1
    assert np.array_equal(got, oracle_render.render(s.words, s.dimension, s.box_a, s.box_b, size))
    # back on the interpreter: same picture
    assert not prog.use_specialized(False)
    assert np.array_equal(image.render_pixels(scene, size), got)


@pytest.mark.parametrize("options", [1, 2])
def test_specialised_ray_caster_of_an_assembly_with_options(cb, scenes, options):
    """The specialised ray caster on the headline assembly with the false-colour and zebra options: the false-colour
    picture counts the steps of every ray, so it would show a single evaluation that went differently."""
    from codecad_b200.cl_util.buffer import ProgramBuffer
    from codecad_b200.rendering import ray_caster
    s = scenes["cfg_planetary"]
    scene = s.compiled()
    scene.program_buffer().specialize(1, ProgramBuffer.SINK_RAY)
    size = (257, 190)
    cam = ray_caster.get_camera_params(scene.bounding_box(), size, None)
    got = ray_caster.render(scene, size=size, options=ray_caster.RenderOptions(options), *cam)
    want = oracle_render.ray_cast(s.words, s.box_a, s.box_b, size, options=options)
    assert np.array_equal(got, want)


def test_specialised_segmented_ray_caster(cb, scenes):
    """The 500-box scene is generated as several functions; one ray per thread goes through them too."""
    from codecad_b200.cl_util.buffer import ProgramBuffer
    from codecad_b200.rendering import image
    s = scenes["cfg_synthetic500"]
    scene = s.compiled()
    scene.program_buffer().specialize(1, ProgramBuffer.SINK_RAY)
    size = (96, 64)
    assert np.array_equal(image.render_pixels(scene, size), oracle_render.ray_cast(s.words, s.box_a, s.box_b, size))


def test_tiered_rendering(cb, scenes):
    """Default mode: the first picture comes from the interpreter while the specialised kernel
    compiles in the background; later pictures use it.  Identical bytes."""
    from codecad_b200 import _lib
    from codecad_b200.cl_util.buffer import ProgramBuffer
    from codecad_b200.rendering import image
    L = _lib.lib()
    s = scenes["dsdf3d_csg_thing"]
    scene = s.compiled()
    size = (160, 120)
    want = oracle_render.render(s.words, 3, s.box_a, s.box_b, size)
    try:
        _lib.check(L.cc_set_jit_mode(1))
        assert np.array_equal(image.render_pixels(scene, size), want)
        ready, _ = scene.program_buffer().wait_specialized(ProgramBuffer.SINK_RAY)
        assert ready == 1
        assert np.array_equal(image.render_pixels(scene, size), want)
    finally:
        _lib.check(L.cc_set_jit_mode(0))


# ---- the operator boundary: opencl_manager.k.<kernel> with cl_util.Buffer, the way the reference's
#      rendering/ray_caster.py:55-75, bitmap.py:27-30 and polygon2d.py:88-117 launch them --------------

def test_kernel_proxy_ray_caster_and_bitmap(cb, scenes):
    from codecad_b200 import cl_util
    from codecad_b200.cl_util import opencl_manager
    from codecad_b200.geometry import BoundingBox, Vector
    from codecad_b200.nodes import make_program_buffer
    from codecad_b200.rendering import ray_caster
    s = scenes["dsdf3d_csg_thing"]
    scene = s.compiled()
    size = (120, 90)
    box = scene.bounding_box()
    origin, direction, up, focal = ray_caster.get_camera_params(box, size, None)
    origin, forward, up, right, tol, radius, dmin, dmax, floor_z = ray_caster.kernel_arguments(box, origin, direction, up, focal)
    out = cl_util.Buffer(np.uint8, [size[0], size[1], 3])
    ev = opencl_manager.k.ray_caster(size, None, make_program_buffer(scene), origin.as_float4(), forward.as_float4(),
                                     up.as_float4(), right.as_float4(), np.float32(tol), np.float32(radius),
                                     np.float32(dmin), np.float32(dmax), np.float32(floor_z), np.uint32(0), out, None)
    got = out.read(wait_for=[ev]).transpose((1, 0, 2))
    assert np.array_equal(got, oracle_render.ray_cast(s.words, s.box_a, s.box_b, size))

    s2 = scenes["dsdf2d_gear"]
    scene2 = s2.compiled()
    box2 = scene2.bounding_box().flattened()
    resolution = Vector(size[0], size[1], 1)
    step = box2.size().elementwise_div(resolution).max()
    o2 = box2.midpoint() - resolution * step / 2
    out2 = cl_util.Buffer(np.uint8, [size[0], size[1], 3])
    ev = opencl_manager.k.bitmap(size, None, make_program_buffer(scene2), o2.as_float4(), np.float32(step), out2)
    got2 = out2.read(wait_for=[ev]).reshape((size[0], size[1], 3)).transpose((1, 0, 2))
    assert np.array_equal(got2, oracle_render.bitmap(s2.words, s2.box_a, s2.box_b, size))


def test_kernel_proxy_process_polygon(cb, scenes):
    import oracle
    from codecad_b200 import cl_util
    from codecad_b200.cl_util import opencl_manager
    from codecad_b200.geometry import FLOAT4, Vector
    from codecad_b200.nodes import make_program_buffer
    s = scenes["dsdf2d_gear"]
    gx = gy = 40
    corner, step = s.grid(40)
    c = Vector(float(corner[0]), float(corner[1]), float(corner[2]))
    corners = cl_util.Buffer(FLOAT4, [gx, gy])
    tri = (gx - 1, gy - 1, 2)
    vertices = cl_util.Buffer(cl_util.Buffer.dual_dtype(np.float32), tri[0] * tri[1] * tri[2])
    links = cl_util.Buffer(np.uint32, tri[0] * tri[1] * tri[2])
    starts = cl_util.Buffer(np.uint32, tri[0] + tri[1])
    counter = cl_util.Buffer(np.uint32, 1)
    prog = make_program_buffer(s.compiled())
    ev1 = opencl_manager.k.grid_eval((gx, gy), None, prog, c.as_float4(), np.float32(step), corners)
    ev2 = counter.enqueue_write(np.zeros(1, np.uint32))
    ev3 = vertices.enqueue_write(np.zeros(tri[0] * tri[1] * tri[2], vertices.dtype))
    ev = opencl_manager.k.process_polygon(tri, None, c.as_float2(), np.float32(step), corners, vertices, links, starts,
                                          counter, wait_for=[ev1, ev2, ev3])
    field = oracle.grid_eval(s.words, corner, step, (gx, gy, 1))[:, :, 0, :]
    want_v, want_l, want_s = oracle.process_polygon(corner[:2], step, field)
    got_v = vertices.read(wait_for=[ev])
    assert np.array_equal(np.stack([got_v["x"], got_v["y"]], -1), want_v)
    assert np.array_equal(links.read(wait_for=[ev]), want_l)
    assert int(counter.read(wait_for=[ev])[0]) == len(want_s)
    assert np.array_equal(starts.read(wait_for=[ev])[:len(want_s)], want_s)


@pytest.mark.parametrize("name", ["dsdf2d_gear", "dsdf2d_nonconvex_shell1", "dsdf3d_csg_thing"])
def test_matplotlib_slice_field(cb, scenes, name):
    """matplotlib_slice.cl:1-20 through rendering.matplotlib_slice.slice_values and through k.matplotlib_slice."""
    import oracle
    from codecad_b200 import cl_util
    from codecad_b200.cl_util import opencl_manager
    from codecad_b200.rendering import matplotlib_slice
    s = scenes[name]
    scene = s.compiled()
    values, corner, resolution, _ = matplotlib_slice.slice_values(scene)
    h, w, three = values.shape
    assert three == 3
    c32 = np.array([corner.x, corner.y, corner.z], np.float64).astype(np.float32)
    field = oracle.grid_eval(s.words, c32, np.float32(resolution), (w, h, 1))[:, :, 0, :]      # [x][y][4]
    want = np.stack([field[..., 3], field[..., 0], field[..., 1]], -1).transpose((1, 0, 2))      # [y][x][3]
    assert np.array_equal(values, want, equal_nan=True)
    out = cl_util.Buffer(np.float32, [h, w, 3])
    ev = opencl_manager.k.matplotlib_slice((w, h), None, scene.program_buffer(), corner.as_float4(), np.float32(resolution), out)
    assert np.array_equal(out.read(wait_for=[ev]).reshape(h, w, 3), want, equal_nan=True)
