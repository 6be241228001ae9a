"""Drop-in boundary against an unmodified reference checkout (this container only: the GPU
box has no /root/reference).  Everything up to the device boundary: the reference imports
with our modules in place of cl_util / grid_eval / subdivision / mass_properties, its own
node compiler produces exactly the committed fixture words, and a compute call without a
GPU fails loudly instead of falling back."""
import os
import random
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("CODECAD_REFERENCE", "/root/reference")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "codecad")),
                                reason="reference checkout not present")


def _run(code):
    env = dict(os.environ)
    # tools/refstub only supplies the `flags` (py-flags) package the image lacks; our own
    # pyopencl stand-in is installed by dropin and takes precedence in sys.modules
    env["PYTHONPATH"] = os.pathsep.join([ROOT, os.path.join(ROOT, "tests"), REF, os.path.join(ROOT, "tools", "refstub")])
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], capture_output=True, text=True, env=env,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_reference_imports_with_our_modules():
    out = _run("""
        import codecad_b200.dropin as d
        codecad = d.load()
        import pyopencl
        print(codecad.cl_util.__name__, codecad.subdivision.__name__, codecad.grid_eval.__name__)
        print(type(codecad.cl_util.opencl_manager).__module__, pyopencl.__doc__[:12])
        print(codecad.mass_properties.__module__)
        print(codecad.cl_util.opencl_manager.max_register_count)
        import codecad.rendering.mesh as m, codecad.rendering.stl_renderer as st
        print(m.triangular_mesh.__module__, st.render_stl.__module__)
    """)
    assert out.strip().splitlines()[-1] == "codecad_b200.rendering.mesh codecad_b200.rendering.stl_renderer"
    lines = out.strip().splitlines()[-5:-1]
    assert lines[0] == "codecad_b200.cl_util codecad_b200.subdivision codecad_b200.grid_eval"
    assert lines[1] == "codecad_b200.cl_util.manager codecad_b200"
    assert lines[2] == "codecad_b200.mass_properties"
    assert lines[3] == "512"


def test_reference_compiler_reproduces_fixture_words():
    out = _run("""
        import random, sys, numpy as np
        import codecad_b200.dropin as d
        codecad = d.load()
        from scenes import load_scenes
        S = load_scenes()
        import codecad.shapes as s
        shapes = {"mp_unit_box": s.box(1), "dsdf3d_torus": s.circle(d=4).translated_x(3).revolved(),
                  "mp_drunk_box": s.box(2, 3, 5).rotated((7, 11, 13), 17),
                  "x_gear3d": s.gears.InvoluteGear(13, 1.0).extruded(2)}
        for name, shape in shapes.items():
            random.seed(0)
            w = codecad.nodes.make_program(shape)
            assert w.dtype == np.float32 and np.array_equal(w, S[name].words), name
            box = shape.bounding_box()
            assert tuple(box.a) == S[name].box_a and tuple(box.b) == S[name].box_b
        print("ok")
    """)
    assert out.strip().endswith("ok")


def test_compute_without_gpu_fails_loudly():
    out = _run("""
        import codecad_b200.dropin as d
        from codecad_b200 import _lib
        codecad = d.load()
        if _lib.load().cc_device_count() > 0:
            print("gpu present")
        else:
            for call in (lambda: codecad.mass_properties(codecad.shapes.box(1), 0.02),
                         lambda: codecad.subdivision.subdivision(codecad.shapes.box(10), 1, grid_size=4),
                         lambda: codecad.nodes.make_program_buffer(codecad.shapes.sphere())):
                try:
                    call()
                except _lib.CodecadB200Error as e:
                    assert "no CPU fallback" in str(e)
                else:
                    raise SystemExit("a compute call succeeded without a GPU")
            print("ok")
    """)
    assert out.strip().splitlines()[-1] in ("ok", "gpu present")


def test_opt_in_deterministic_program_build():
    """make_program(shape, random_passes=0): reproducible, much faster, same geometry."""
    out = _run("""
        import time, numpy as np
        import codecad_b200.dropin as d
        codecad = d.load()
        import codecad.shapes as s
        from codecad_b200.nodes import make_program
        from codecad_b200 import _lib
        import oracle
        parts = [s.box(3, 4, 5).rotated((1, 2, 3), 10 * i).translated(4 * i, 0, 0) for i in range(12)]
        shape = s.union(parts) - s.sphere(3).translated(6, 1, 0)
        t0 = time.perf_counter(); w_ref = make_program(shape); t_ref = time.perf_counter() - t0
        t0 = time.perf_counter(); w0 = make_program(shape, random_passes=0); t_fast = time.perf_counter() - t0
        assert np.array_equal(w0, make_program(shape, random_passes=0))          # reproducible
        assert t_fast * 5 < t_ref, (t_fast, t_ref)
        info, _ = _lib.decode_program(w0)                                          # a valid program
        assert info.n_instructions > 40
        corner, step, dims = np.array([-6, -8, -8], np.float32), np.float32(0.9), (60, 18, 18)
        a = oracle.grid_eval(w_ref, corner, step, dims)
        b = oracle.grid_eval(w0, corner, step, dims)
        assert np.allclose(a[..., 3], b[..., 3], rtol=1e-5, atol=1e-5)             # same distances
        print("ok %.2f s -> %.3f s" % (t_ref, t_fast))
    """)
    assert out.strip().splitlines()[-1].startswith("ok")
