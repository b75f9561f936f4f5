"""Block-level breakdown of an `ncu --page source --csv` dump: consecutive SASS instructions with
(nearly) equal execution counts are merged; prints share of executed warp instructions, active
threads per instruction and the opcode mix.   python scripts/sass_blocks.py X_source.csv [min_share]"""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
hi = [i for i, r in enumerate(rows) if 'Source' in r and any('Instructions Executed' in c for c in r)][0]
hdr = rows[hi]; ci = hdr.index('Source')
ie = [i for i, c in enumerate(hdr) if c.strip() == 'Instructions Executed'][0]
te = [i for i, c in enumerate(hdr) if 'Thread Instructions Executed' in c][0]
data = []
for r in rows[hi + 1:]:
    try:
        v = int(r[ie]); t = int(r[te])
    except ValueError:
        continue
    data.append((r[0], r[ci].strip(), v, t))
tot = sum(d[2] for d in data)
blocks = []
for a, s, v, t in data:
    op = s.split()[1] if s.startswith('@') else s.split()[0]
    if blocks and v > 0 and abs(blocks[-1]['v'] - v) <= 0.02 * v:
        b = blocks[-1]; b['n'] += 1; b['sum'] += v; b['t'] += t; b['ops'].append(op)
    else:
        blocks.append({'a': a, 'v': v, 'n': 1, 'sum': v, 't': t, 'ops': [op]})
print("total executed warp instructions", tot)
for b in blocks:
    if b['sum'] / tot > minshare:
        c = Counter(o.split('.')[0] for o in b['ops'])
        print(f"{100 * b['sum'] / tot:5.1f}%  n={b['n']:4d} exec/instr={b['v']:.3g} thr={b['t'] / max(b['sum'], 1):4.1f}  {dict(c.most_common(7))}")
