"""Parity at BASELINE.json's full sizes, through properties that do not need the oracle to
evaluate 10^9 points: whole x-planes of the full grid against the oracle, tier 1 == tier 2 on a
sample of planes (checksum of checksums), slab-sharded == unsharded, and the hierarchy configs
end to end against the oracle's host algorithms."""
import zlib

import numpy as np
import pytest

import oracle
from oracle import host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import codecad_b200
    from codecad_b200 import _lib
    _lib.init(0)
    old = _lib.check(_lib.lib().cc_set_jit_mode(0))
    yield codecad_b200
    _lib.check(_lib.lib().cc_set_jit_mode(old))


def _planes(L, buf, n, xs):
    """Copy whole x-planes of a device-resident n^3 float4 grid to the host."""
    import ctypes
    from codecad_b200 import _lib
    out = {}
    plane_bytes = n * n * 16
    for x in xs:
        host_arr = np.empty((n, n, 4), np.float32)
        _lib.check(L.cc_memcpy_d2h_async(host_arr.ctypes.data, ctypes.c_void_p(buf.device_ptr.value + x * plane_bytes),
                                         plane_bytes, None))
        _lib.check(L.cc_synchronize())
        out[x] = host_arr
    return out


def test_planetary_1024_cubed(cb, scenes):
    """configs[3]: dense 1024^3 grid_eval of the planetary gearbox, the bench workload."""
    from codecad_b200 import _lib
    from codecad_b200.cl_util import Buffer
    from codecad_b200.cl_util.buffer import ProgramBuffer
    from codecad_b200.geometry import FLOAT4
    L = _lib.lib()
    s = scenes["cfg_planetary"]
    n = 1024
    corner, step = s.grid(n)
    prog = ProgramBuffer(s.words)
    out = Buffer(FLOAT4, (n, n, n))
    rng = np.random.default_rng(1024)
    sample = sorted(int(x) for x in rng.choice(n, size=48, replace=False))
    # tier 1 (interpreter)
    cb.grid_eval(prog, corner, step, (n, n, n), device_out=out)
    t1 = _planes(L, out, n, sample)
    # whole planes against the oracle (bit-exact)
    for x in sample[:3]:
        want = oracle.grid_eval(s.words, corner, step, (1, n, n), x_offset=x)[0]
        assert np.array_equal(t1[x], want, equal_nan=True), "plane %d differs from the oracle" % x
    crc1 = zlib.crc32(b"".join(np.uint32(zlib.crc32(t1[x].tobytes())).tobytes() for x in sample))
    # tier 2 (specialised kernel): same checksum of checksums
    prog.specialize(0, ProgramBuffer.SINK_FLOAT4)
    _lib.check(L.cc_memset_async(out.device_ptr, 0, n * n * n * 16, None))
    cb.grid_eval(prog, corner, step, (n, n, n), device_out=out)
    t2 = _planes(L, out, n, sample)
    crc2 = zlib.crc32(b"".join(np.uint32(zlib.crc32(t2[x].tobytes())).tobytes() for x in sample))
    assert crc1 == crc2
    # tier 3 (the part-culling pair, what dense grids of this assembly use in steady state): the same planes,
    # and every byte of a 128-plane slab through the middle of the gearbox against tier 2
    mid = Buffer(FLOAT4, (128, n, n))
    cb.grid_eval(prog, corner, step, (128, n, n), x_offset=448, device_out=mid)
    want_mid = np.empty((128, n, n, 4), np.float32)
    _lib.check(L.cc_memcpy_d2h_async(want_mid.ctypes.data, mid.device_ptr, want_mid.nbytes, None))
    _lib.check(L.cc_synchronize())
    prog.specialize(0, ProgramBuffer.SINK_PARTS)
    launches0, _ = _lib.counters()
    _lib.check(L.cc_memset_async(out.device_ptr, 0, n * n * n * 16, None))
    cb.grid_eval(prog, corner, step, (n, n, n), device_out=out)
    assert _lib.counters()[0] == launches0 + 2, "the part-culling kernels did not run"
    t3 = _planes(L, out, n, sample)
    crc3 = zlib.crc32(b"".join(np.uint32(zlib.crc32(t3[x].tobytes())).tobytes() for x in sample))
    assert crc1 == crc3
    _lib.check(L.cc_memset_async(mid.device_ptr, 0, want_mid.nbytes, None))
    cb.grid_eval(prog, corner, step, (128, n, n), x_offset=448, device_out=mid)
    got_mid = np.empty_like(want_mid)
    _lib.check(L.cc_memcpy_d2h_async(got_mid.ctypes.data, mid.device_ptr, got_mid.nbytes, None))
    _lib.check(L.cc_synchronize())
    assert got_mid.tobytes() == want_mid.tobytes()
    # tier 4 (the column kernels on top of the part masks: the gears' 2-D profiles once per z-column): same again
    prog.specialize(0, ProgramBuffer.SINK_COLUMNS)
    launches0, _ = _lib.counters()
    _lib.check(L.cc_memset_async(out.device_ptr, 0, n * n * n * 16, None))
    cb.grid_eval(prog, corner, step, (n, n, n), device_out=out)
    assert _lib.counters()[0] == launches0 + 4, "the column kernels did not run (centres, column pass, bricks, flagged bricks)"
    t4 = _planes(L, out, n, sample)
    crc4 = zlib.crc32(b"".join(np.uint32(zlib.crc32(t4[x].tobytes())).tobytes() for x in sample))
    assert crc1 == crc4
    _lib.check(L.cc_memset_async(mid.device_ptr, 0, want_mid.nbytes, None))
    cb.grid_eval(prog, corner, step, (128, n, n), x_offset=448, device_out=mid)
    _lib.check(L.cc_memcpy_d2h_async(got_mid.ctypes.data, mid.device_ptr, got_mid.nbytes, None))
    _lib.check(L.cc_synchronize())
    assert got_mid.tobytes() == want_mid.tobytes()
    mid.release()
    del got_mid, want_mid
    # z-slab sharding as on 8 GPUs: rank 5 of 8 evaluates x in [640, 768) with an offset
    x0, x1 = cb.grid_eval.__globals__["slab_range"](n, 5, 8)
    slab = Buffer(FLOAT4, (x1 - x0, n, n))
    cb.grid_eval(prog, corner, step, (x1 - x0, n, n), x_offset=x0, device_out=slab)
    inside = [x for x in sample if x0 <= x < x1] or [x0]
    if inside == [x0]:
        t2.update(_planes(L, out, n, [x0]))
    got = _planes(L, slab, n, [x - x0 for x in inside])
    for x in inside:
        assert np.array_equal(got[x - x0], t2[x], equal_nan=True)
    out.release()
    slab.release()


def test_synthetic500_2048_slab(cb, scenes):
    """configs[4]: the 500-box scene on the 2048^3 grid; one GPU of eight owns 256 x-planes.
    A strip of this rank's slab against the oracle, and the slab offset against offset 0 geometry."""
    s = scenes["cfg_synthetic500"]
    n = 2048
    corner, step = s.grid(n)
    x0, _ = cb.grid_eval.__globals__["slab_range"](n, 3, 8)
    dims = (2, 24, n)
    got = cb.grid_eval(s.compiled(), corner, step, dims, x_offset=x0 + 100)
    got = np.stack([got["x"], got["y"], got["z"], got["w"]], axis=-1)
    want = oracle.grid_eval(s.words, corner, step, dims, x_offset=x0 + 100)
    assert np.array_equal(got, want, equal_nan=True)


def test_synthetic500_culled_slab_equals_the_full_walk(cb, scenes):
    """64 x-planes of the 2048^3 grid (2.7e8 points) of the 500-box scene: the union-forest kernel against
    the interpreter, which evaluates all 500 boxes at every point — every byte."""
    from codecad_b200 import _lib
    L = _lib.lib()
    s = scenes["cfg_synthetic500"]
    n = 2048
    corner, step = s.grid(n)
    scene = s.compiled()
    dims = (64, n, n)
    _lib.check(L.cc_set_forest_mode(1))
    got = np.array(cb.grid_eval(scene, corner, step, dims, x_offset=1000))
    try:
        _lib.check(L.cc_set_forest_mode(0))
        want = np.array(cb.grid_eval(scene, corner, step, dims, x_offset=1000))
    finally:
        _lib.check(L.cc_set_forest_mode(1))
    assert got.tobytes() == want.tobytes()


def test_airfoil_mass_properties_config(cb, scenes):
    """configs[2]: examples/airfoil.py mass_properties to 1e-6 relative (here: 1e-12 against the
    oracle's restatement of the reference's host loop, at three resolutions)."""
    a = scenes["cfg_airfoil"]
    for res in (1.0, 0.5, 0.25):
        vol, cen, inertia = host.mass_properties(a.words, a.box_a, a.box_b, res, 64)
        got = cb.mass_properties(a.compiled(), res, 64)
        assert got.volume == pytest.approx(vol, rel=1e-12)
        assert np.allclose(got.centroid, cen, rtol=1e-12, atol=1e-12)
        assert np.allclose(got.inertia_tensor, inertia, rtol=1e-9, atol=1e-9 * np.abs(inertia).max())
    # the same call through the specialised kernels, and through the column kernels (the two 163-edge profiles once
    # per column of every 64^3 block): integer sums, so the same doubles to the last bit
    from codecad_b200 import _lib
    from codecad_b200.cl_util.buffer import ProgramBuffer
    scene = a.compiled()
    interp = cb.mass_properties(scene, 0.25, 64)
    scene.program_buffer().specialize(0, ProgramBuffer.SINK_MASS)
    spec = cb.mass_properties(scene, 0.25, 64)
    scene.program_buffer().specialize(0, ProgramBuffer.SINK_TILES_MASS)
    n0 = _lib.counters()[0]
    cols = cb.mass_properties(scene, 0.25, 64)
    n_cols = _lib.counters()[0] - n0
    old = _lib.check(_lib.lib().cc_set_columns_mode(0))
    n0 = _lib.counters()[0]
    cb.mass_properties(scene, 0.25, 64)
    assert n_cols > _lib.counters()[0] - n0, "the column pass did not run"
    _lib.check(_lib.lib().cc_set_columns_mode(old))
    for other in (spec, cols):
        assert other.volume == interp.volume and tuple(other.centroid) == tuple(interp.centroid)
        assert np.array_equal(np.array(other.inertia_tensor), np.array(interp.inertia_tensor))


@pytest.mark.parametrize("grid", [128, 16])
def test_csg_subdivision_config(cb, scenes, grid):
    """configs[1]: csg_example at 512^3 effective resolution: the leaf blocks are the oracle's."""
    c = scenes["cfg_csg_example"]
    res = 100.0 / 512
    dims, want = host.subdivision(c.words, c.box_a, c.box_b, 3, res, True, grid)
    _, got_dims, got = cb.subdivision(c.compiled(), res, True, grid)
    assert tuple(got_dims) == tuple(dims)
    assert len(got) == {128: 112, 16: 11936}[grid]
    assert sorted((tuple(b[3]), tuple(b[1])) for b in got) == sorted((tuple(b[3]), tuple(b[1])) for b in want)


def test_menger_256_against_oracle_planes(cb, scenes):
    """configs[0]: menger sponge at 256^3 (the reference's CPU-runnable case)."""
    s = scenes["cfg_menger_sponge"]
    n = 256
    corner, step = s.grid(n)
    got = cb.grid_eval(s.compiled(), corner, step, (n, n, n))
    got = np.stack([got["x"], got["y"], got["z"], got["w"]], axis=-1)
    for x in (0, 77, 128, 255):
        want = oracle.grid_eval(s.words, corner, step, (1, n, n), x_offset=x)[0]
        assert np.array_equal(got[x], want, equal_nan=True)


@pytest.mark.parametrize("name", ["cfg_planetary", "cfg_menger_sponge", "cfg_airfoil", "cfg_csg_example"])
def test_config_scenes_rendered_at_the_reference_size(cb, scenes, name):
    """The reference's default picture, 1024x768 (rendering/image.py:7), of the BASELINE scenes: both
    tiers byte-identical to the oracle's restatement of the ray caster (~11 M evaluate() calls)."""
    from oracle import render as oracle_render
    from codecad_b200.cl_util.buffer import ProgramBuffer
    from codecad_b200.rendering import image
    s = scenes[name]
    scene = s.compiled()
    size = (1024, 768)
    want = oracle_render.ray_cast(s.words, s.box_a, s.box_b, size)
    assert np.array_equal(image.render_pixels(scene, size), want)
    scene.program_buffer().specialize(1, ProgramBuffer.SINK_RAY)
    assert scene.program_buffer().use_specialized(True)
    assert np.array_equal(image.render_pixels(scene, size), want)
    assert (want != want[0, 0]).any()          # not a blank picture
