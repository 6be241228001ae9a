"""STL export — same entry point as /root/reference/codecad/rendering/stl_renderer.py:8-24
`render_stl(obj, filename)`.  The reference fills a numpy-stl Mesh triangle by triangle in Python
and lets numpy-stl (1.8.0, requirements.txt:12, not vendored) write a binary STL; here the binary
STL is written directly from the triangle soup (80-byte header, uint32 count, then per triangle
normal + 3 vertices as float32 + a zero attribute word)."""
import numpy as np

from .mesh import mesh_arrays


def write_binary_stl(filename, triangles, header=b"codecad_b200"):
    """triangles: float [t][3][3]."""
    tri = np.asarray(triangles, dtype=np.float32).reshape(-1, 3, 3)
    rec = np.zeros(len(tri), dtype=np.dtype([("normal", "<f4", 3), ("v", "<f4", (3, 3)), ("attr", "<u2")]))
    rec["v"] = tri
    # numpy-stl's update_normals(): cross(v1 - v0, v2 - v0), not normalised
    rec["normal"] = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    with open(filename, "wb") as f:
        f.write(header[:80].ljust(80, b" "))
        f.write(np.uint32(len(tri)).tobytes())
        f.write(rec.tobytes())
    return len(tri)


def render_stl(obj, filename):
    vertices, _, _ = mesh_arrays(obj)
    return write_binary_stl(filename, vertices)
