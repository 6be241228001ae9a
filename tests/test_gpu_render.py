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

@pytest.mark.parametrize("name", NAMES + ["cfg_planetary", "cfg_menger_sponge"])
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
