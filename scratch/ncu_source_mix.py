#!/usr/bin/env python
"""Instruction mix and execution plateaus of one kernel from an ncu report's source page.
    python scratch/ncu_source_mix.py REP.ncu-rep KERNEL_REGEX [instance]"""
import collections
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# several kernel instances follow each other: take the first block
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []
        blocks.append(cur)
    elif cur is not None:
        cur.append(r)
blk = blocks[int(sys.argv[3]) if len(sys.argv) > 3 else 0]
hdr, data = blk[0], [r for r in blk[1:] if len(r) == len(blk[0])]
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[iE]) for r in data)
print("total warp instructions", tot, "static", len(data))
mix, samp = collections.Counter(), collections.Counter()
for r in data:
    t = r[iS].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    mix[op] += int(r[iE])
    samp[op] += int(r[iSamp])
for op, c in mix.most_common(28):
    print(f"{op:10s} {c / 1e6:8.1f}M  {100 * c / tot:5.1f}%  samples {samp[op]}")
print("-- plateaus (index, executed, instruction)")
last = None
for idx, r in enumerate(data):
    e = int(r[iE])
    if last is None or abs(e - last) > 0.25 * max(e, last, 1):
        print(idx, e, r[iS][:80])
    last = e
