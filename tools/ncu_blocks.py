#!/usr/bin/env python
"""Group an `ncu --page source --csv` (SASS view) export into straight-line blocks of equal execution count and print
each block's share of executed instructions, stall samples, shared-memory wavefronts and opcode mix."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
h = rows[1]
ie = h.index('Instructions Executed'); src = h.index('Source'); ws = h.index('L1 Wavefronts Shared'); wi = h.index('L1 Wavefronts Shared Ideal'); smp = h.index('# Samples')
data = []; seen = set()
for r in rows[2:]:
    if len(r) <= ie or r[0] in seen: continue
    seen.add(r[0])
    try: n = int(r[ie])
    except ValueError: continue
    data.append((r[0], n, r[src], float(r[ws] or 0), float(r[wi] or 0), int(r[smp] or 0)))
tot = sum(d[1] for d in data); tots = sum(d[5] for d in data)
print('total warp-instructions', tot, 'samples', tots)
blocks = []; cur = []; prev = None
for d in data:
    if prev is not None and abs(d[1] - prev) > 0.15 * max(d[1], prev, 1):
        blocks.append(cur); cur = []
    cur.append(d); prev = d[1]
blocks.append(cur)
for b in blocks:
    n = sum(d[1] for d in b); s = sum(d[5] for d in b)
    if n < tot * thr and s < tots * thr: continue
    ops = collections.Counter(d[2].split()[0] if not d[2].startswith('@') else d[2].split()[1] for d in b)
    w = sum(d[3] for d in b); wid = sum(d[4] for d in b)
    print("%s..%s insts=%3d exec/inst=%9d share=%5.1f%% samples=%5.1f%% wf=%d ideal=%d  %s" % (b[0][0][-5:], b[-1][0][-5:], len(b), b[0][1], 100 * n / tot, 100 * s / tots, w, wid, dict(ops.most_common(6))))
