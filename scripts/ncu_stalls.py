"""Top stall sites of one kernel from an ncu report's SASS source page:
   ncu -i X.ncu-rep --page source --csv --print-source sass --kernel-id :::K > sass.csv ; python scripts/ncu_stalls.py sass.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
data = []
for r in rows[hi + 1:]:                      # first kernel section only
    if r and r[0] == 'Kernel Name': break
    if len(r) == len(hdr): data.append(r)
ia = hdr.index('# Samples'); isrc = hdr.index('Source'); iex = hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ia]) for r in data)
print('kernel', rows[0][1][:110] if rows[0] else '', '\ntotal samples', tot, 'instructions', len(data))
agg = {}
for r in data:
    for c in stall_cols:
        agg[hdr[c][6:]] = agg.get(hdr[c][6:], 0) + int(r[c])
print('stall totals', sorted(agg.items(), key=lambda kv: -kv[1])[:8])
top = sorted(range(len(data)), key=lambda i: -int(data[i][ia]))[:topn]
for i in sorted(top):
    r = data[i]
    st = sorted(((hdr[c][6:], int(r[c])) for c in stall_cols if int(r[c]) > 0), key=lambda kv: -kv[1])[:3]
    print('%5d %6s %8s  %-72s %s' % (i, r[ia], r[iex], r[isrc].strip()[:72], st))
