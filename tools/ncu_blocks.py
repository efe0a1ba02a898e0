"""Summarise an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv) into hot SASS basic blocks."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
hdr = rows[1]; data = rows[2:]
isrc = hdr.index('Source'); ie = hdr.index('Instructions Executed'); ismp = hdr.index('# Samples')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ie]) for r in data); tots = sum(int(r[ismp]) for r in data)
print('total inst', tot, 'samples', tots, 'n sass', len(data))
blocks = []; cur = None
for i, r in enumerate(data):
    e = int(r[ie]); s = int(r[ismp])
    if cur and cur['e'] == e: cur['n'] += 1; cur['s'] += s; cur['end'] = i
    else:
        cur = dict(start=i, end=i, e=e, n=1, s=s); blocks.append(cur)
for b in blocks:
    w = b['e'] * b['n']
    if w / tot > thr or b['s'] / tots > thr:
        ops = {}; st = {}
        for r in data[b['start']:b['end'] + 1]:
            t = r[isrc].split()
            op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
            ops[op] = ops.get(op, 0) + 1
            for c in stall_cols:
                v = int(r[c] or 0)
                if v: st[hdr[c]] = st.get(hdr[c], 0) + v
        top = sorted(ops.items(), key=lambda x: -x[1])[:7]
        tst = sorted(st.items(), key=lambda x: -x[1])[:3]
        print(f"[{b['start']:4d}-{b['end']:4d}] n={b['n']:4d} exec={b['e']:9d} inst%={100*w/tot:5.1f} samp%={100*b['s']/tots:5.1f} {top} {tst}")
