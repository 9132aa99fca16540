#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv`, long format: one row per launch and metric).

usage: python profiles/launch_summary.py launches.csv [updates_per_step] [algorithmic_bytes_per_update] [--one-step]

--one-step: keep only the launches of ONE complete evolution() (from the launch after a k_bnd_macros up to and
including the next k_bnd_macros, the last kernel of a step).
"""
import collections
import csv
import sys

one_step = "--one-step" in sys.argv
if one_step:
    sys.argv.remove("--one-step")
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
if one_step:
    ends = sorted({int(r[0]) for r in rows if r[4].startswith("k_bnd_macros")})
    # the first k_bnd_macros belongs to dugks_create (the launches up to the second one include the initialisation
    # kernels): take the window between the last two
    if len(ends) >= 3:
        rows = [r for r in rows if ends[-2] < int(r[0]) <= ends[-1]]
    elif len(ends) == 2:
        rows = [r for r in rows if ends[0] < int(r[0]) <= ends[1]]
per = collections.defaultdict(dict)
name = {}
for r in rows:
    per[int(r[0])][r[12]] = float(r[14].replace(",", ""))
    name[int(r[0])] = r[4]
unit_t = next((r[13] for r in rows if r[12] == "gpu__time_duration.sum"), "ns")
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit_t, 1e-3)
agg = collections.defaultdict(lambda: [0.0, 0, 0.0, 0.0])
for i, m in per.items():
    a = agg[name[i]]
    a[0] += m.get("gpu__time_duration.sum", 0.0) * scale
    a[1] += 1
    a[2] += m.get("dram__bytes_read.sum", 0.0)
    a[3] += m.get("dram__bytes_write.sum", 0.0)
tot = sum(a[0] for a in agg.values())
print(f"launches {len(per)} total {tot:.0f} us (ncu: cold-cache, serialised launches; compare shares)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{a[0]:10.0f} us {a[1]:4d} {a[0] / tot * 100:5.1f}% {a[0] / a[1]:9.1f} us/launch  dram R {a[2] / a[1] / 1e9:6.2f} "
          f"W {a[3] / a[1] / 1e9:6.2f} GB/launch  {k[:70]}")
dram = sum(a[2] + a[3] for a in agg.values())
if len(sys.argv) > 2:
    upd = float(sys.argv[2])
    alg = f" (algorithmic {float(sys.argv[3]):.2f})" if len(sys.argv) > 3 else ""
    print(f"DRAM bytes of these launches {dram / 1e9:.1f} GB = {dram / upd:.1f} B per update{alg}")
else:
    print(f"DRAM bytes of these launches {dram / 1e9:.1f} GB")
