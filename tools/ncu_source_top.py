#!/usr/bin/env python3
"""Summarise `ncu --page source --csv` of one kernel: stall samples per reason, and the SASS lines with the most samples.
usage: ncu -i rep.ncu-rep --page source --csv | tools/ncu_source_top.py [N]"""
import csv
import sys

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rows = list(csv.reader(sys.stdin))
h, data = rows[1], rows[2:]
col = {k: i for i, k in enumerate(h)}
samp = [int(r[col["# Samples"]]) for r in data]
tot = sum(samp)
print("kernel:", rows[0][1])
print("samples %d, warp instructions %d, SASS lines %d" % (tot, sum(int(r[col["Instructions Executed"]]) for r in data), len(data)))
reasons = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(int(r[col[k]]) for r in data) for k in reasons}
print("stall samples:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 200 > tot))
ops = {}
for r, s in zip(data, samp):
    op = r[col["Source"]].split()
    op = (op[1] if op and op[0].startswith("@") else op[0]) if op else "?"
    op = op.split(".")[0]
    e = ops.setdefault(op, [0, 0])
    e[0] += s
    e[1] += int(r[col["Instructions Executed"]])
print("by opcode (samples %, warp instructions):", ", ".join("%s %.1f%% %d" % (k, 100.0 * v[0] / tot, v[1]) for k, v in sorted(ops.items(), key=lambda kv: -kv[1][0])[:14]))
print("top lines (index, samples, top stall, executed, SASS):")
for i in sorted(sorted(range(len(data)), key=lambda i: -samp[i])[:n]):
    r = data[i]
    top = max(reasons, key=lambda k: int(r[col[k]]))
    print("%5d %6d %-12s %9s  %s" % (i, samp[i], top[6:], r[col["Instructions Executed"]], r[col["Source"]].strip()[:100]))
