"""Two-GPU path (skipped on a one-GPU box): x-slab sharded grid_eval needs no collective and
reproduces the unsharded grid bit for bit; sharded mass_properties all-reduces ten float64
integrals over NCCL and every rank returns the single-GPU result."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import importlib
    import codecad_b200
    from codecad_b200 import _lib
    from scenes import load_scenes
    _lib.init(rank)
    ge = importlib.import_module("codecad_b200.grid_eval")
    S = load_scenes()
    s = S["cfg_planetary"]
    dims = (37, 16, 32)
    corner, step = s.grid(40)
    x0, x1 = ge.slab_range(dims[0], rank, world)
    part = codecad_b200.grid_eval(s.compiled(), corner, step, (x1 - x0, dims[1], dims[2]), x_offset=x0)
    np.save(os.path.join(out_dir, "grid_r%d.npy" % rank), np.array(part))
    a = S["cfg_airfoil"]
    mp = codecad_b200.mass_properties(a.compiled(), 0.5, 64, group=True)
    np.save(os.path.join(out_dir, "mp_r%d.npy" % rank),
            np.concatenate([[mp.volume], list(mp.centroid), mp.inertia_tensor.ravel()]))
    blocks = codecad_b200.subdivision(S["cfg_csg_example"].compiled(), 100 / 128, True, 16, rank=rank, world=world)[2]
    np.save(os.path.join(out_dir, "sub_r%d.npy" % rank), np.array([b[3] for b in blocks], dtype=np.int64).reshape(-1, 3))
    dist.destroy_process_group()


def test_two_gpu_sharding(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)

    import codecad_b200
    from codecad_b200 import _lib
    from scenes import load_scenes
    _lib.init(0)
    S = load_scenes()
    s = S["cfg_planetary"]
    dims = (37, 16, 32)
    corner, step = s.grid(40)
    full = np.array(codecad_b200.grid_eval(s.compiled(), corner, step, dims))
    parts = np.concatenate([np.load(tmp_path / ("grid_r%d.npy" % r)) for r in range(2)], axis=0)
    assert parts.tobytes() == full.tobytes()

    single = codecad_b200.mass_properties(S["cfg_airfoil"].compiled(), 0.5, 64)
    want = np.concatenate([[single.volume], list(single.centroid), single.inertia_tensor.ravel()])
    r0, r1 = np.load(tmp_path / "mp_r0.npy"), np.load(tmp_path / "mp_r1.npy")
    assert np.array_equal(r0, r1)
    assert np.array_equal(r0, want), "the int64 all-reduce of the exact accumulator is bit-identical to one GPU"

    whole = codecad_b200.subdivision(S["cfg_csg_example"].compiled(), 100 / 128, True, 16)[2]
    got = np.concatenate([np.load(tmp_path / ("sub_r%d.npy" % r)) for r in range(2)])
    assert sorted(map(tuple, got.tolist())) == sorted(tuple(b[3]) for b in whole)


SINGLE_PROCESS = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import codecad_b200
from codecad_b200 import _lib
from scenes import load_scenes
S = load_scenes()
air, csg, plan = S["cfg_airfoil"].compiled(), S["cfg_csg_example"].compiled(), S["cfg_planetary"]
_lib.init(0)
one_mp = codecad_b200.mass_properties(air, 0.5, 64)
one_sub = codecad_b200.subdivision(csg, 100 / 512, True, 16)[2]
corner, step = plan.grid(48)
one_grid = np.array(codecad_b200.grid_eval(plan.compiled(), corner, step, (45, 16, 32)))
_lib.init_devices(list(range(%(n)d)))          # the same process now drives every GPU
assert _lib.active_devices() == %(n)d
st = {}
all_mp = codecad_b200.mass_properties(air, 0.5, 64, stats=st)
all_sub = codecad_b200.subdivision(csg, 100 / 512, True, 16)[2]
all_grid = np.array(codecad_b200.grid_eval(plan.compiled(), corner, step, (45, 16, 32)))
assert st["devices"] == %(n)d and st["cells_idlest_device"] > 0, st
assert all_mp.volume == one_mp.volume and tuple(all_mp.centroid) == tuple(one_mp.centroid)
assert np.array_equal(all_mp.inertia_tensor, one_mp.inertia_tensor)
assert np.array_equal(all_sub.int_corners, one_sub.int_corners) and np.array_equal(all_sub.corners, one_sub.corners)
assert all_grid.tobytes() == one_grid.tobytes()
print("single-process ok", st)
"""


def test_one_process_drives_all_gpus(tmp_path):
    """cc_init_devices: the unmodified module calls use every GPU of the box from ONE process and
    return what one GPU returns, bit for bit."""
    import subprocess
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    r = subprocess.run([sys.executable, "-c", SINGLE_PROCESS % {"root": ROOT, "n": n}], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "single-process ok" in r.stdout
