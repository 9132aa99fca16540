#!/usr/bin/env python
"""Executed instructions and stall samples per SOURCE LINE of one kernel: joins the SASS-level source page of an ncu
report (`ncu -i rep --page source --csv`, as profiles/ncu_export.sh writes it) with the line table of the cubin
(`cuobjdump -xelf all libdugks.so; nvdisasm --print-line-info -c *.cubin`; kernels are built with -lineinfo).

usage: python profiles/line_profile.py <ncu source csv> <nvdisasm listing> <mangled kernel name> [top N]
"""
import collections
import csv
import re
import sys

src_csv, sass, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
txt = open(sass).read().split("\n")
start = [i for i, l in enumerate(txt) if l.startswith("//--------------------- .text." + kname)][0]
lines, cur = [], None
for l in txt[start + 1:]:
    if l.startswith("//--------------------- ."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    elif re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
        lines.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
iN, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
body = [r for r in rows[2:] if len(r) > iS]
assert len(body) == len(lines), (len(body), len(lines))
cnt, smp = collections.Counter(), collections.Counter()
for s, r in zip(lines, body):
    cnt[s] += int(r[iN] or 0)
    smp[s] += int(r[iS] or 0)
tot, ts = sum(cnt.values()), sum(smp.values())
print(f"{kname}: {tot} warp instructions, {ts} stall samples")
for s, n in cnt.most_common(top):
    print(f"  {s[0]}:{s[1]:<5d} {n / tot * 100:5.1f}% of instructions  {smp[s] / ts * 100:5.1f}% of samples")
