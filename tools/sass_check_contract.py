#!/usr/bin/env python3
"""Static check of a cubin: is any FADD2 fed directly by an FMUL2 (the pattern ptxas contracts), and do FMUL2s
remain whose consumers are not FFMA2?  Linear scan per function, no control-flow analysis (conservative hint)."""
import re, subprocess, sys
sass = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
last = {}
n_fmul2 = n_fadd2 = flagged = 0
for line in sass.splitlines():
    if "Function :" in line:
        last = {}
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s+(.*?);", line)
    if not m:
        continue
    addr, op, args = m.groups()
    regs = re.findall(r"\bR(\d+)", args)
    if not regs:
        continue
    dst, srcs = int(regs[0]), [int(r) for r in regs[1:]]
    base = op.split(".")[0]
    if base == "FADD2":
        n_fadd2 += 1
        for s in srcs:
            w = last.get(s)
            if w and w[0] == "FMUL2":
                flagged += 1
                print("FADD2 at %s reads R%d written by FMUL2 at %s" % (addr, s, w[1]))
    if base == "FMUL2":
        n_fmul2 += 1
    wide = base in ("FADD2", "FMUL2", "FFMA2") or ".64" in op
    last[dst] = (base, addr)
    if wide:
        last[dst + 1] = (base, addr)
print("FMUL2 %d, FADD2 %d, FADD2-fed-by-FMUL2 %d" % (n_fmul2, n_fadd2, flagged))
