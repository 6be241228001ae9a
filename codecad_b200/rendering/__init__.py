"""Callers of the hot path that the reference keeps under codecad/rendering and that were widened
into: mesh export.  Everything else in the reference's rendering package (ray
caster, 2-D outlines, matplotlib viewers, CLI) is out of scope."""
from .mesh import triangular_mesh, mesh_arrays  # noqa: F401
from .stl_renderer import render_stl, write_binary_stl  # noqa: F401
