"""Volume / centroid / inertia tensor — same signature and result type as the
reference's /root/reference/codecad/mass_properties.py:30-229.  The per-block job loop
(:69-159: one launch, three blocking reads and a Python generator per 64^3 block) is
replaced by cc_mass_properties: every level is one batched launch over all of its blocks,
hit lists and the float64 corner chain stay on the device, and only ten float64 integrals
come back (summed over ranks with an NCCL all-reduce when sharded).
"""
import collections
import ctypes

import numpy as np

from . import _lib
from .geometry import Vector, as_vector
from .nodes import make_program_buffer
from .subdivision import _levels, calculate_block_sizes


class MassProperties(collections.namedtuple("MassProperties", "volume centroid inertia_tensor")):
    """Volume, centroid and inertia tensor (referenced to the centroid, not the origin)."""

    __slots__ = ()


STAT_NAMES = ("launches", "cells", "blocks", "levels", "dealt_blocks", "devices", "cells_busiest_device",
              "cells_idlest_device")


def mass_limbs(program, box_a, resolution, block_sizes, rank=0, world=1):
    """This rank's share of the ten integrals as the library's exact accumulator: (limbs int64[40],
    exponents int32[10], stats dict).  Limbs of different ranks add up exactly (integers);
    limbs_to_integrals() converts the total."""
    limbs = (ctypes.c_int64 * 40)()
    exps = (ctypes.c_int32 * 10)()
    stats = (ctypes.c_uint64 * 8)()
    a = (ctypes.c_double * 3)(float(box_a[0]), float(box_a[1]), float(box_a[2]))
    _lib.check(_lib.lib().cc_mass_properties_exact(program.handle, a, float(resolution), _levels(block_sizes),
                                                   len(block_sizes), int(rank), int(world), limbs, exps, stats))
    return (np.array(limbs[:], dtype=np.int64), np.array(exps[:], dtype=np.int32),
            dict(zip(STAT_NAMES, (int(v) for v in stats))))


def limbs_to_integrals(limbs, exponents):
    """Exact accumulator -> the ten float64 integrals (one rounding each)."""
    l = (ctypes.c_int64 * 40)(*[int(v) for v in limbs])
    e = (ctypes.c_int32 * 10)(*[int(v) for v in exponents])
    out = (ctypes.c_double * 10)()
    _lib.check(_lib.load().cc_mass_limbs_to_integrals(l, e, out))
    return np.array(out[:], dtype=np.float64)


def mass_integrals(program, box_a, resolution, block_sizes, rank=0, world=1):
    """The ten integrals (one,x,y,z,xx,yy,zz,xy,xz,yz) over this rank's share, + stats
    (launches, cells, blocks, levels)."""
    limbs, exps, stats = mass_limbs(program, box_a, resolution, block_sizes, rank, world)
    return limbs_to_integrals(limbs, exps), tuple(stats[k] for k in STAT_NAMES[:4])


def finish(integrals):
    """mass_properties.py:179-229: integrals -> MassProperties."""
    (integral_one, integral_x, integral_y, integral_z, integral_xx, integral_yy, integral_zz,
     integral_xy, integral_xz, integral_yz) = (float(v) for v in integrals)

    volume = integral_one
    if volume == 0:
        return MassProperties(0, Vector.splat(0), np.zeros((3, 3)))
    centroid = Vector(integral_x, integral_y, integral_z) / integral_one

    sxx = integral_xx - 2 * centroid.x * integral_x + centroid.x * centroid.x * integral_one
    syy = integral_yy - 2 * centroid.y * integral_y + centroid.y * centroid.y * integral_one
    szz = integral_zz - 2 * centroid.z * integral_z + centroid.z * centroid.z * integral_one
    sxy = integral_xy - centroid.x * integral_y - centroid.y * integral_x + centroid.x * centroid.y * integral_one
    sxz = integral_xz - centroid.x * integral_z - centroid.z * integral_x + centroid.x * centroid.z * integral_one
    syz = integral_yz - centroid.y * integral_z - centroid.z * integral_y + centroid.y * centroid.z * integral_one
    I_xx, I_yy, I_zz = syy + szz, sxx + szz, sxx + syy
    I_xy, I_xz, I_yz = -sxy, -sxz, -syz
    inertia_tensor = np.array([[I_xx, I_xy, I_xz], [I_xy, I_yy, I_yz], [I_xz, I_yz, I_zz]])
    return MassProperties(volume, centroid, inertia_tensor)


def mass_properties(shape, resolution, grid_size=None, group=None, stats=None):
    """mass_properties.py:30.  With `group` (a torch.distributed process group, or True
    for the default group) the hierarchy is sharded over its ranks and the exact accumulator
    of the ten integrals is all-reduced (40 int64): every rank returns the same result, and it
    is bit-identical to the unsharded one.  One process that drives several GPUs
    (_lib.init_devices / CODECAD_B200_DEVICES) shards inside the library with no argument here.
    `stats` (a dict) receives this rank's counters."""
    if grid_size is None:
        grid_size = 64

    assert shape.dimension() == 3, "2D objects are not supported yet"
    assert resolution > 0, "Non-positive resolution makes no sense"
    assert grid_size > 1, "Grid needs to be at least 2x2x2"
    assert grid_size ** 5 <= 2 ** 32, "Centroid coordinate sums would overflow"

    program_buffer = make_program_buffer(shape)
    bb = shape.bounding_box()
    box_a, box_b = as_vector(bb[0]), as_vector(bb[1])
    block_sizes = calculate_block_sizes((box_a, box_b), 3, resolution, grid_size, overlap=False)

    rank, world = 0, 1
    if group is not None:
        import torch.distributed as dist
        pg = None if group is True else group
        rank, world = dist.get_rank(pg), dist.get_world_size(pg)

    limbs, exps, st = mass_limbs(program_buffer, box_a, resolution, block_sizes, rank, world)
    if stats is not None:
        stats.update(st)
    if world > 1:
        limbs = allreduce_limbs(limbs, None if group is True else group)
    return finish(limbs_to_integrals(limbs, exps))


def allreduce_limbs(limbs, group=None):
    """Sum the exact accumulators (40 int64) over the ranks — the only data that crosses NVLink on
    this path.  Integer sums are exact and order-independent, so every rank converts the same total.
    NCCL needs the tensor on the GPU; gloo (CPU tests) does not."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(limbs, dtype=np.int64))
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def allreduce_integrals(integrals, group=None):
    """Sum the ten float64 partial integrals over the ranks (the only data that crosses
    NVLink on this path).  NCCL needs the tensor on the GPU; gloo (CPU tests) does not."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(integrals, dtype=np.float64))
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()
