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
    assert np.allclose(r0, want, rtol=1e-12, atol=0)

    whole = codecad_b200.subdivision(S["cfg_csg_example"].compiled(), 100 / 128, True, 16)[2]
    got = np.concatenate([np.load(tmp_path / ("sub_r%d.npy" % r)) for r in range(2)])
    assert sorted(map(tuple, got.tolist())) == sorted(tuple(b[3]) for b in whole)
