"""N>1 host path on CPU: world_size-2 gloo process group.  The only data that crosses ranks
on this path is the ten float64 mass-property integrals (mass_properties.allreduce_integrals);
grids and subdivision blocks are partitioned with no collective."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib
    mpm = importlib.import_module("codecad_b200.mass_properties")
    from codecad_b200.grid_eval import slab_range
    # each rank contributes the integrals of "its" half of a unit cube split at x = 0.5
    x0, x1 = (0.0, 0.5) if rank == 0 else (0.5, 1.0)
    vol = (x1 - x0)
    ints = np.array([vol, vol * (x0 + x1) / 2, vol * 0.5, vol * 0.5,
                     (x1 ** 3 - x0 ** 3) / 3, vol / 3, vol / 3,
                     vol * (x0 + x1) / 2 * 0.5, vol * (x0 + x1) / 2 * 0.5, vol * 0.25])
    total = mpm.allreduce_integrals(ints)
    res = mpm.finish(total)
    slabs = [slab_range(1024, r, world) for r in range(world)]
    np.save(os.path.join(out_dir, "r%d.npy" % rank),
            np.concatenate([[res.volume], list(res.centroid), np.diag(res.inertia_tensor), np.ravel(slabs)]))
    dist.destroy_process_group()


def test_two_rank_allreduce_of_integrals(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "r0.npy")
    r1 = np.load(tmp_path / "r1.npy")
    assert np.array_equal(r0, r1), "every rank must return the same result"
    assert r0[0] == pytest.approx(1.0)
    assert tuple(r0[1:4]) == pytest.approx((0.5, 0.5, 0.5))
    assert tuple(r0[4:7]) == pytest.approx((1 / 6, 1 / 6, 1 / 6))     # unit cube about its centroid
    assert list(r0[7:]) == [0, 512, 512, 1024]
