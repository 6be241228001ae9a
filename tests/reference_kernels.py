"""Host-side counterparts of the kernels the reference's OWN tests bring along (tests/test_dsdf.cl,
tests/test_clutil.cl), registered with the replacement opencl_manager so that those tests run
unmodified against the CUDA path.  Test infrastructure: every evaluate() goes to the GPU through
codecad_b200.evaluate_points (cc_evaluate_points); what stays here is the few lines of glue each
test kernel has around it."""
import numpy as np


def _points(global_size, corner, step):
    dims = tuple(int(v) for v in global_size) + (1,) * (3 - len(global_size))
    c = np.asarray(corner)
    c = np.array([c["x"], c["y"], c["z"]], dtype=np.float32) if c.dtype.names else np.asarray(c, np.float32)[:3]
    idx = np.stack(np.meshgrid(*[np.arange(d, dtype=np.float32) for d in dims], indexing="ij"), axis=-1)
    return dims, (c + np.float32(step) * idx).astype(np.float32).reshape(-1, 3)   # boxCorner + boxStep * coords


class _Done:
    def wait(self):
        return self


def register(manager, evaluate_words):
    """evaluate_words(program_buffer, points[n][3]) -> float32 [n][4]"""

    def grid_eval_twice(global_size, local_size, scene, corner, step, output, wait_for=None):   # test_dsdf.cl:26-40
        dims, p = _points(global_size, corner, step)
        first = evaluate_words(scene, p)
        second = evaluate_words(scene, (p - first[:, :3] * first[:, 3:4]).astype(np.float32))
        arr = output._process_array(None)
        both = np.stack([first, second], axis=1).reshape(dims + (2, 4))
        for k, name in enumerate("xyzw"):
            arr[name] = both[..., k]
        output.enqueue_write().wait()
        return _Done()

    def estimate_direction(global_size, local_size, scene, corner, step, epsilon, output, wait_for=None):  # :7-24
        dims, p = _points(global_size, corner, step)
        eps = np.float32(epsilon)
        centre = evaluate_words(scene, p)[:, 3]
        plus = np.stack([evaluate_words(scene, (p - eps * e).astype(np.float32))[:, 3] for e in np.eye(3, dtype=np.float32)], 1)
        minus = np.stack([evaluate_words(scene, (p + eps * e).astype(np.float32))[:, 3] for e in np.eye(3, dtype=np.float32)], 1)
        arr = output._process_array(None)
        both = np.stack([centre[:, None] - plus, minus - centre[:, None]], axis=1).reshape(dims + (2, 3))
        for k, name in enumerate("xyz"):
            arr[name] = both[..., k]
        output.enqueue_write().wait()
        return _Done()

    def actual_distance_to_surface(global_size, local_size, step, inp, output, wait_for=None):  # :42-70
        dims = tuple(int(v) for v in global_size) + (1,) * (3 - len(global_size))
        w = inp.read()["w"][..., 0].reshape(-1)
        sign = np.sign(w)
        idx = np.stack(np.meshgrid(*[np.arange(d, dtype=np.float32) for d in dims], indexing="ij"), axis=-1).reshape(-1, 3)
        pts = np.float32(step) * idx
        out = np.empty(len(pts), np.float32)
        for s0 in (-1.0, 0.0, 1.0):
            mine = np.where(sign == s0)[0]
            other = pts[sign != s0]
            if len(mine) == 0:
                continue
            if len(other) == 0:
                out[mine] = np.sqrt(np.float32(np.finfo(np.float32).max))
                continue
            for a in range(0, len(mine), 512):
                q = pts[mine[a:a + 512]]
                d2 = ((q[:, None, :] - other[None, :, :]) ** 2).sum(-1)
                out[mine[a:a + 512]] = np.sqrt(d2.min(axis=1))
        output._process_array(None)[...] = out.reshape(dims)
        output.enqueue_write().wait()
        return _Done()

    # tests/test_clutil.cl: the two kernels that only exercise Buffer plumbing
    def one_item_double(global_size, local_size, value, wait_for=None):
        a = value.read()
        a[...] = a * 2
        value.enqueue_write().wait()
        return _Done()

    def indexing_identity(global_size, local_size, output, wait_for=None):
        dims = tuple(int(v) for v in global_size)
        arr = output._process_array(None)
        full = dims + (1,) * (3 - len(dims))
        grid = np.meshgrid(*[np.arange(d, dtype=np.uint32) for d in full], indexing="ij")
        for k, name in enumerate("xyz"):
            arr[name] = grid[k].reshape(arr.shape)
        output.enqueue_write().wait()
        return _Done()

    for fn in (grid_eval_twice, estimate_direction, actual_distance_to_surface, one_item_double, indexing_identity):
        manager.register_kernel(fn.__name__, fn)
