"""CPU: the oracle's restatement of the 2-D outline path (rendering/polygon2d.cl + polygon2d.py) against
known answers.  The reference has no test of its own for this path; the properties below are the ones
its consumers (rendering/svg.py) rely on: closed outlines, vertices on the surface, the right area."""
import math

import numpy as np
import pytest

import oracle
from oracle import host
from scenes import ALL_NAMES

NAMES_2D = [n for n in ALL_NAMES if n.startswith("dsdf2d_")]


def signed_area(chain):
    a = np.asarray(chain, np.float64)
    x, y = a[:, 0], a[:, 1]
    return 0.5 * float(np.sum(x * np.roll(y, -1) - y * np.roll(x, -1)))


def canonical(chains):
    """Chains as a sorted list of tuples, each rotated to start at its smallest vertex."""
    out = []
    for c in chains:
        c = [tuple(v) for v in c]
        i = c.index(min(c))
        out.append(tuple(c[i:] + c[:i]))
    return sorted(out)


@pytest.mark.parametrize("name,expected,outlines", [
    ("dsdf2d_circle", math.pi * 4, 1), ("dsdf2d_rectangle", 8.0, 1), ("dsdf2d_polygon2d_square", 25.0, 1),
    ("dsdf2d_polygon2d_triangle", 3.0, 1)])
def test_area_and_outline_count(scenes, name, expected, outlines):
    s = scenes[name]
    for grid in (128, 16, 7):
        chains = host.polygon(s.words, s.box_a, s.box_b, s.feature_size / 8, grid)
        assert len(chains) == outlines
        assert sum(signed_area(c) for c in chains) == pytest.approx(expected, rel=5e-3)  # place_vertex stops at residual^2 < 1e-3


@pytest.mark.parametrize("name", NAMES_2D)
def test_vertices_lie_on_the_surface(scenes, name):
    s = scenes[name]
    resolution = s.feature_size / 8          # finer than the reference's default feature_size / 2
    chains = host.polygon(s.words, s.box_a, s.box_b, s.feature_size / 4, 128)
    assert chains
    for c in chains:
        assert len(c) >= 3
        pts = np.zeros((len(c), 3), np.float32)
        pts[:, :2] = np.asarray(c, np.float32)
        d = oracle.evaluate_points(s.words, pts)[:, 3]
        # a vertex sits in a triangle the surface crosses; the plane fit moves it onto the surface
        # except at sharp corners, where it stays within the cell
        assert np.abs(d).max() <= resolution * 1.5
        # (the search stops once the summed squared residual of the three planes is < 1e-3)
        assert np.median(np.abs(d)) <= max(resolution * 0.25, math.sqrt(1e-3))


@pytest.mark.parametrize("name", ["dsdf2d_gear", "dsdf2d_nonconvex_shell1", "dsdf2d_bin_counter_11", "dsdf2d_mirror_2d"])
def test_box_size_does_not_change_the_outlines(scenes, name):
    """Small boxes force outlines across many box borders: same count, same enclosed area."""
    s = scenes[name]
    ref = host.polygon(s.words, s.box_a, s.box_b, s.feature_size, 128)
    for grid in (24, 9, 5):
        chains = host.polygon(s.words, s.box_a, s.box_b, s.feature_size, grid)
        assert len(chains) == len(ref)
        assert sorted(round(signed_area(c), 3) for c in chains) == pytest.approx(
            sorted(round(signed_area(c), 3) for c in ref), abs=0.02 * s.feature_size ** 2 * 10)
        # winding: material on a fixed side everywhere -> the total keeps its sign and value
        assert sum(signed_area(c) for c in chains) == pytest.approx(sum(signed_area(c) for c in ref), rel=5e-3)


def test_process_polygon_link_encoding():
    """One box, a vertical edge at x = 1.5: every surface triangle links to the next one up or
    down, the chain enters and leaves through the box border (polygon2d.cl:5-35 encoding)."""
    gx = gy = 4
    xs = np.arange(gx, dtype=np.float32)
    field = np.zeros((gx, gy, 4), np.float32)
    field[..., 0] = 1.0                       # gradient +x
    field[..., 3] = (xs - 1.5)[:, None]       # distance: inside for x < 1.5
    vertices, links, starts = oracle.process_polygon((0.0, 0.0), 1.0, field)
    assert len(starts) == 1
    surface = np.flatnonzero((links & 0xFFF00000) == 0)
    out = np.flatnonzero((links & 0x80000000) != 0) if False else [i for i in range(len(links)) if links[i] != 0xFFFFFFFF]
    assert len(out) == 6                      # 2 triangles per row x 3 rows of cells crossed
    assert len(surface) == 5                  # all but the one that leaves the box
    assert np.allclose(vertices[out][:, 0], 1.5, atol=1e-6)
    start = int(starts[0])
    assert start & 0x80000000 and (start & 0xFFFFF) in out


@pytest.mark.parametrize("grid", [128, 24, 9, 5])
@pytest.mark.parametrize("name", ["dsdf2d_gear", "dsdf2d_nonconvex_shell1", "dsdf2d_bin_counter_11", "dsdf2d_circle"])
def test_product_host_logic_on_oracle_data(scenes, name, grid, monkeypatch):
    """The product's chain following / piece joining (codecad_b200/rendering/polygon2d.py), fed with the
    oracle's per-box outputs instead of the device's: same polygons as the oracle's own procedure."""
    from codecad_b200.rendering import polygon2d
    s = scenes[name]

    def fake_subdivision(obj, resolution, grid_size=None):
        max_dims, boxes = host.subdivision(s.words, s.box_a, s.box_b, 2, resolution, True, grid_size)
        return None, max_dims, boxes

    def fake_polygon_blocks(program_buffer, grid_xy, corners, resolution):
        gx, gy = grid_xy
        cells = 2 * (gx - 1) * (gy - 1)
        n = len(corners)
        vertices = np.zeros((n, cells, 2), np.float32)
        links = np.zeros((n, cells), np.uint32)
        starts = np.zeros((n, gx + gy - 2), np.uint32)
        counts = np.zeros(n, np.uint32)
        for b, c in enumerate(corners):
            c32 = np.asarray(c, np.float64).astype(np.float32)
            field = oracle.grid_eval(s.words, c32, resolution, (gx, gy, 1))[:, :, 0, :]
            vertices[b], links[b], st = oracle.process_polygon(c32, resolution, field)
            starts[b, :len(st)] = st
            counts[b] = len(st)
        return vertices, links, starts, counts

    monkeypatch.setattr(polygon2d, "subdivision", fake_subdivision)
    monkeypatch.setattr(polygon2d, "polygon_blocks", fake_polygon_blocks)
    got = list(polygon2d.polygon(s.compiled(), grid))
    want = host.polygon(s.words, s.box_a, s.box_b, s.feature_size, grid)
    assert canonical(got) == canonical(want)
    if grid == 128:
        assert [[tuple(v) for v in c] for c in got] == [[tuple(v) for v in c] for c in want]


def test_open_outline_raises_like_the_reference():
    """A surface that leaves the only box: the reference ends on `assert len(open_chain_beginnings) == 0`
    (polygon2d.py:172-173); the native assembly reports CC_ERR_OPEN_OUTLINE -> AssertionError."""
    from codecad_b200.rendering import polygon2d
    gx = gy = 4
    xs = np.arange(gx, dtype=np.float32)
    field = np.zeros((gx, gy, 4), np.float32)
    field[..., 0] = 1.0
    field[..., 3] = (xs - 1.5)[:, None]
    vertices, links, starts = oracle.process_polygon((0.0, 0.0), 1.0, field)
    st = np.zeros((1, gx + gy - 2), np.uint32)
    st[0, :len(starts)] = starts
    with pytest.raises(AssertionError):
        polygon2d.assemble(vertices[None], links[None].copy(), st, np.array([len(starts)], np.uint32), [[0, 0, 0]], 3)
    # a closed square inside one box assembles without a device
    d = np.maximum(np.abs(xs[:, None] - 1.5), np.abs(xs[None, :] - 1.5)) - 1.0
    field2 = np.zeros((gx, gy, 4), np.float32)
    field2[..., 3] = d
    field2[..., 0] = np.where(np.abs(xs[:, None] - 1.5) >= np.abs(xs[None, :] - 1.5), np.sign(xs[:, None] - 1.5), 0)
    field2[..., 1] = np.where(np.abs(xs[:, None] - 1.5) < np.abs(xs[None, :] - 1.5), np.sign(xs[None, :] - 1.5), 0)
    v2, l2, s2 = oracle.process_polygon((0.0, 0.0), 1.0, field2)
    assert len(s2) == 0
    chains = polygon2d.assemble(v2[None], l2[None].copy(), np.zeros((1, 6), np.uint32), np.zeros(1, np.uint32), [[0, 0, 0]], 3)
    assert len(chains) == 1 and len(chains[0]) >= 4
    assert abs(signed_area(chains[0])) == pytest.approx(4.0, rel=0.2)
