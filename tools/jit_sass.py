#!/usr/bin/env python3
"""Generate the scene-specialised source for a fixture scene, compile it with nvcc (same flags as the
NVRTC path) and print the SASS instruction mix of the float4 kernel:  scene pts [threads] [minb]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
name, pts = sys.argv[1], int(sys.argv[2])
if len(sys.argv) > 3:
    os.environ["CODECAD_B200_JIT_THREADS"] = sys.argv[3]
if len(sys.argv) > 4:
    os.environ["CODECAD_B200_JIT_MINB"] = sys.argv[4]
from codecad_b200 import _lib  # noqa: E402
from scenes import load_scenes  # noqa: E402

src = _lib.specialize_source(load_scenes()[name].words, pts, compile=False, sink_mask=1)
src = src[0] if isinstance(src, tuple) else src
os.makedirs("/tmp/jit", exist_ok=True)
open("/tmp/jit/scene.cu", "w").write(src)
csrc = os.path.join(ROOT, "codecad_b200", "csrc")
out = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false", "-std=c++17", "-I", csrc,
                      "-cubin", "-Xptxas", "-v", "-o", "/tmp/jit/scene.cubin", "/tmp/jit/scene.cu"],
                     capture_output=True, text=True)
print("\n".join(l for l in out.stderr.splitlines() if "registers" in l or "error" in l or "spill" in l))
sass = subprocess.run(["cuobjdump", "-sass", "/tmp/jit/scene.cubin"], capture_output=True, text=True).stdout
cnt = collections.Counter()
fn = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and fn == "cc_jit_float4":
        cnt[m.group(2).split(".")[0]] += 1
tot = sum(cnt.values())
print("cc_jit_float4: %d SASS instructions (static, main function only)" % tot)
print(", ".join("%s %d" % kv for kv in cnt.most_common(25)))
