#!/usr/bin/env python3
"""usage: ncu_hot.py <src.csv.gz> : opcode histogram (executed warp instructions) and stall samples
of one kernel from `ncu --page source --csv --print-source sass`"""
import csv, gzip, sys, collections, re
rows = list(csv.reader(gzip.open(sys.argv[1], "rt")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
cols = {n: i for i, n in enumerate(rows[hdr])}
ops, samples = collections.Counter(), collections.Counter()
tot = 0
for r in rows[hdr + 1:]:
    if len(r) < len(cols): continue
    src = r[cols["Source"]].strip()
    m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", src)
    op = m.group(2) if m else src
    n = int(r[cols["Instructions Executed"]] or 0)
    ops[op] += n; tot += n
    samples[op] += int(r[cols["# Samples"]] or 0)
print("total warp instructions", tot, "static", len(rows) - hdr - 1)
st = sum(samples.values())
for op, n in ops.most_common(25):
    print(f"{op:10s} {n:14d} {100*n/tot:5.1f}%  samples {100*samples[op]/max(st,1):5.1f}%")
