#!/usr/bin/env python
"""Bucket the SASS instructions of an .ncu-rep by how often they execute per work item.
usage: python scripts/ncu_buckets.py rep.ncu-rep <warp-level executions of one per-item instruction>"""
import collections, csv, io, subprocess, sys
rep, unit = sys.argv[1], float(sys.argv[2])
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]; ix = {c: i for i, c in enumerate(h)}
b = collections.defaultdict(collections.Counter)
tot = collections.Counter()
for r in rows[2:]:
    try:
        n = float(r[ix['Instructions Executed']])
    except Exception:
        continue
    op = r[ix['Source']].split()
    op = op[1] if op[0].startswith('@') else op[0]
    op = op.split('.')[0]
    k = round(n / unit, 2)
    b[k][op] += 1
    tot[k] += 1
for k in sorted(b, key=lambda k: -k * tot[k])[:14]:
    print('x%-7s static %4d  dyn/item %8.1f  %s' % (k, tot[k], k * tot[k], dict(b[k].most_common(14))))
