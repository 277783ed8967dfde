#!/usr/bin/env python
"""Aggregates `ncu --page source --csv` output by SASS opcode: executed warp instructions and stall samples.
Usage: ncu -i report.ncu-rep --page source --csv --kernel-name regex:<name> > src.csv; python tools/ncu_opcode_mix.py src.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
header_index = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[header_index]
ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
ops, samples = collections.Counter(), collections.Counter()
total = static = 0
for r in rows[header_index + 1:]:
    if len(r) <= ia or not r[ia].isdigit():
        continue
    src = r[isrc].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2).split(".")[0] if m else src[:10]
    n = int(r[ia])
    ops[op] += n
    samples[op] += int(r[isamp]) if r[isamp].isdigit() else 0
    total += n
    static += 1
print("total warp instructions", total, "static SASS instructions", static)
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print("%-10s %12d %5.1f%%  stall samples %d" % (op, n, 100.0 * n / total, samples[op]))
