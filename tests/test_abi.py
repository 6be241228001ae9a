"""The C-ABI shared library loads, exports exactly the symbols include/codecad_b200.h
declares, and refuses to work without a CUDA device (no fallback)."""
import ctypes
import os
import re

import pytest

from codecad_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "codecad_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    L = _lib.load()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "libcodecad_b200.so does not export %s" % n
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding and header disagree"


def test_no_torch_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "codecad_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)      # declarations only, comments stripped
    assert "torch" not in code and "at::" not in code and "Tensor" not in code
    assert re.findall(r"#include\s*<([^>]+)>", code) == ["stddef.h", "stdint.h"]


def test_fails_loudly_without_a_gpu():
    L = _lib.load()
    if L.cc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    assert L.cc_init(0) < 0
    assert b"no CPU fallback" in L.cc_last_error()
    # every compute entry point refuses to run
    assert L.cc_synchronize() == -2          # CC_ERR_NOT_INITIALIZED
    p = ctypes.c_void_p()
    import numpy as np
    w = np.zeros(4, np.float32)
    assert L.cc_program_create(w.ctypes.data_as(_lib.c_float_p), 4, ctypes.byref(p)) == -2
    with pytest.raises(_lib.CodecadB200Error):
        _lib.init(0)


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under codecad_b200/ may import or load it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "codecad_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M) or "liboracle" in text \
                        or "libcodecad_ref" in text:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
