#!/usr/bin/env python
"""Opcode histogram (executed warp-instructions, stall samples) from `ncu --page source --csv`.
usage: ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME | python profiles/sass_hist.py [topN]"""
import collections
import csv
import sys

rows = list(csv.reader(sys.stdin))
blocks = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
for bi, hi in enumerate(blocks):
    hdr = rows[hi]
    si, ei, wi = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
    end = blocks[bi + 1] - 1 if bi + 1 < len(blocks) else len(rows)
    name = rows[hi - 1][1] if hi > 0 and len(rows[hi - 1]) > 1 else ''
    tot = 0
    ops, stall = collections.Counter(), collections.Counter()
    for r in rows[hi + 1:end]:
        if len(r) <= max(ei, wi) or not r[ei].isdigit():
            continue
        n = int(r[ei]); tot += n
        toks = r[si].split()
        op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '')
        ops[op.split('.')[0]] += n
        stall[op.split('.')[0]] += int(r[wi] or 0)
    print('==', name[:80], 'total warp-instructions', tot)
    top = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    st = sum(stall.values()) or 1
    for k, v in ops.most_common(top):
        print(f"   {k:12s} {v:12d} {v / tot * 100:5.1f}%   stall samples {stall[k] / st * 100:5.1f}%")
