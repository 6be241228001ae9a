#!/usr/bin/env python3
"""Executed-branch algorithmic flop/point of the BASELINE.json scenes: the instrumented
CPU oracle counts, per evaluation, the reference-formulation cost of the branches actually taken
(oracle/sdf_oracle.c FL()), over a 64^3 stratified subsample of each config's grid.  Writes
profiles/executed_flops.json, which bench.py reports beside the static minimum (tests/test_host_logic.py
checks the file against the oracle).   python tools/measure_executed_flops.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from scenes import load_scenes  # noqa: E402

CONFIGS = {  # scene: (grid n, samples per axis)
    "cfg_menger_sponge": (256, 64), "cfg_csg_example": (1024, 64), "cfg_airfoil": (1024, 64),
    "cfg_planetary": (1024, 64), "cfg_synthetic500": (2048, 32),
}


def measure():
    S = load_scenes()
    out = {}
    for name, (n, samples) in CONFIGS.items():
        s = S[name]
        corner, step = s.grid(n)
        mean, pts = oracle.executed_flops(s.words, corner, step, (n, n, n), stride=n // samples)
        out[name] = {"grid": n, "sample_points": pts, "flop_per_point_executed": round(mean, 2)}
    return out


if __name__ == "__main__":
    data = {"source": "tools/measure_executed_flops.py: instrumented CPU oracle, stratified subsample of the config grid, "
                      "SURVEY.md 8(a3) counting rules (reference formulation of every op)",
            "scenes": measure()}
    path = os.path.join(ROOT, "profiles", "executed_flops.json")
    with open(path, "w") as f:
        json.dump(data, f, indent=1)
        f.write("\n")
    print(json.dumps(data, indent=1))
