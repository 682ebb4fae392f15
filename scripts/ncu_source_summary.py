"""Summarise `ncu --page source --csv --print-source sass` output: instruction mix, stall reasons, hottest SASS lines.
Usage: python scripts/ncu_source_summary.py gpurun_out/src.csv [top_n]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) >= len(h)]
ci = {n: i for i, n in enumerate(h)}
ops = collections.Counter(); samp = collections.Counter(); stalls = collections.Counter()
stall_cols = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
ti = ts = 0
for r in data:
    inst = int(r[ci['Instructions Executed']] or 0); s = int(r[ci['# Samples']] or 0)
    toks = r[ci['Source']].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    ops[op] += inst; samp[op] += s; ti += inst; ts += s
    for c in stall_cols:
        stalls[c] += int(r[ci[c]] or 0)
print('total warp inst', ti, 'samples', ts)
print('ops by inst:', ops.most_common(22))
print('ops by samples:', samp.most_common(14))
print('stalls:', stalls.most_common(10))
print('hottest lines (# samples, inst executed, index, sass):')
idx = {id(r): i for i, r in enumerate(data)}
for r in sorted(data, key=lambda r: -int(r[ci['# Samples']] or 0))[:topn]:
    print(f"{r[ci['# Samples']]:>6} {r[ci['Instructions Executed']]:>8} {idx[id(r)]:5d}  {r[ci['Source']][:100]}")
