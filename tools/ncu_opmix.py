#!/usr/bin/env python3
"""Executed-instruction mix by SASS opcode from an ncu report's source page:  tools/ncu_opmix.py report.ncu-rep [top]"""
import collections
import csv
import io
import re
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
while rows and "Source" not in rows[0]:
    rows.pop(0)
hdr = rows[0]
isrc = hdr.index("Source")
iex = hdr.index("Instructions Executed")
cnt = collections.Counter()
tot = 0
for r in rows[1:]:
    if len(r) <= iex:
        continue
    try:
        n = int(float(r[iex]))
    except ValueError:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[isrc])
    if not m:
        continue
    cnt[m.group(2)] += n
    tot += n
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
print("total warp-level instructions executed: %d" % tot)
for k, v in cnt.most_common(top):
    print("  %-12s %14d  %5.1f %%" % (k, v, 100.0 * v / tot))
