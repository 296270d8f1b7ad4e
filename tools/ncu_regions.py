#!/usr/bin/env python
"""Per-region view of ONE kernel from `ncu --import-source on ... ; ncu -i rep --page source --csv`: consecutive SASS instructions
with the same execution count form a region (a loop body or a straight-line stretch); for every region that matters the share of
executed warp instructions, of shared-memory wavefronts (and what they would be without bank conflicts), of global L1 tag requests
and of the warp-state samples.   usage: ncu_regions.py source.csv [min_share_percent]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
floor = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
name = rows[0][1]
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, KeyError, IndexError):
        return 0.0


keys = ("Instructions Executed", "Thread Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "L1 Tag Requests Global", "# Samples")
tot = {k: sum(f(r, k) for r in data) or 1.0 for k in keys}
print(name)
print(f"SASS instructions {len(data)}; executed warp instructions {tot[keys[0]]:.0f}; thread instructions {tot[keys[1]]:.0f}; "
      f"shared-memory wavefronts {tot[keys[2]]:.0f} (ideal {tot[keys[3]]:.0f}); global L1 tag requests {tot[keys[4]]:.0f}; samples {tot[keys[5]]:.0f}")
groups, cur = [], None
for i, r in enumerate(data):
    ne = f(r, keys[0])
    if cur and abs(cur["ne"] - ne) <= 0.02 * max(ne, 1.0):
        cur["end"] = i
    else:
        cur = {"start": i, "end": i, "ne": ne, **{k: 0.0 for k in keys}}
        groups.append(cur)
    for k in keys:
        cur[k] += f(r, k)
print(f"{'SASS rows':>12} {'n':>4} {'exec/instr':>11} {'inst %':>7} {'smem wf %':>9} {'ideal %':>8} {'L1 tag %':>8} {'samples %':>9}  first instruction")
for g in groups:
    if 100 * g[keys[0]] / tot[keys[0]] >= floor:
        print(f"{g['start']:5d}-{g['end']:5d} {g['end'] - g['start'] + 1:4d} {g['ne']:11.0f} {100 * g[keys[0]] / tot[keys[0]]:7.1f} "
              f"{100 * g[keys[2]] / tot[keys[2]]:9.1f} {100 * g[keys[3]] / tot[keys[2]]:8.1f} {100 * g[keys[4]] / tot[keys[4]]:8.1f} "
              f"{100 * g[keys[5]] / tot[keys[5]]:9.1f}  {data[g['start']][ix['Source']].strip()[:48]}")
