#!/usr/bin/env python3
"""FP32 operations a kernel ISSUED, from an `ncu --metrics smsp__sass_thread_inst_executed_op_{ffma,ffma2,
fmul,fmul2,fadd,fadd2}_pred_on.sum ... --csv` log:  tools/ncu_issued_flops.py log.csv points
Thread-level counts; FFMA = 2 flop, a packed instruction (FFMA2 / FMUL2 / FADD2) works on two lanes."""
import csv
import json
import sys


def parse(path):
    vals = {}
    for r in csv.reader(open(path)):
        if len(r) > 14 and r[12].startswith(("smsp__", "dram__", "gpu__")):
            vals[r[12]] = float(r[14])
    return vals


def issued(vals, points):
    g = lambda k: vals.get("smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % k, 0.0)  # noqa: E731
    flops = 2 * g("ffma") + 4 * g("ffma2") + g("fmul") + 2 * g("fmul2") + g("fadd") + 2 * g("fadd2")
    seconds = vals["gpu__time_duration.sum"] * 1e-9
    return {"flop_per_point_issued": flops / points,
            "ffma2_per_point": g("ffma2") / points, "fmul2_per_point": g("fmul2") / points, "fadd2_per_point": g("fadd2") / points,
            "ffma_per_point": g("ffma") / points, "fmul_per_point": g("fmul") / points, "fadd_per_point": g("fadd") / points,
            "fp32_thread_instructions_per_point": vals.get("smsp__sass_thread_inst_executed_op_fp32_pred_on.sum", 0.0) / points,
            "warp_instructions_per_point": vals.get("smsp__inst_executed.sum", 0.0) / points,
            "tflops_issued_under_ncu": flops / seconds / 1e12, "ms_under_ncu": seconds * 1e3,
            "dram_bytes": vals.get("dram__bytes_read.sum", 0.0) + vals.get("dram__bytes_write.sum", 0.0)}


if __name__ == "__main__":
    print(json.dumps(issued(parse(sys.argv[1]), float(sys.argv[2])), indent=1))
