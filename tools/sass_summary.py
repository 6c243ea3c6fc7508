#!/usr/bin/env python3
"""SASS evidence for profiles/: per hot kernel, the opcode histogram of `cuobjdump -sass` of the built library and the lines that show
which hardware paths it uses (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA engine),
SYNCS = mbarrier, LDGSTS = cp.async, HMMA = legacy mma.sync, DFMA / DADD = fp64).
usage: tools/sass_summary.py [kernel-substring ...] > profiles/rN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "art_b200", "libart_hotpath.so")
WANT = sys.argv[1:] or ["k_dn_blocks5", "k_shrink_v", "k_shrink_h", "k_shrink_sf", "k_mad_hist_all", "k_fat_dct", "k_dirinterp", "k_wav_sy_sub", "k_dn_gather"]
MARK = ("UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "SYNCS", "LDGSTS", "HMMA", "FENCE.VIEW.ASYNC")

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        funcs[name] = []
    elif name and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
        funcs[name].append(line)
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
print("# cuobjdump -sass art_b200/libart_hotpath.so (sm_100a), nvcc %s" % subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-2])
for n, lines in funcs.items():
    if not any(w in n for w in WANT):
        continue
    ops = collections.Counter()
    marks = collections.OrderedDict()
    for ln in lines:
        body = ln.split("*/", 1)[1].split("/*")[0].strip().rstrip(";").strip()
        toks = body.split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        ops[op.split(".")[0]] += 1
        for mk in MARK:
            if op.startswith(mk) and len(marks.setdefault(mk, [])) < 3:
                marks[mk].append(re.sub(r"\s+", " ", body))
    print("\n## %s\n   %d SASS instructions" % (demangle(n), len(lines)))
    print("   opcodes: " + ", ".join("%s %d" % kv for kv in ops.most_common(24)))
    for mk, ex in marks.items():
        print("   %-10s x%-4d e.g. %s" % (mk, sum(v for k, v in ops.items() if k == mk.split(".")[0]) or len(ex), " | ".join(ex[:2])))
