"""Summarise `ncu --page source --csv` output: hottest SASS lines of one kernel."""
import csv, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r][0]
hdr = rows[hi]
A, S, IE, WS = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
data = []
for r in rows[hi + 1:]:
    try:
        data.append((int(r[A], 16), float(r[IE] or 0), float(r[WS] or 0), r[S]))
    except Exception:
        pass
tot = sum(d[1] for d in data); tots = sum(d[2] for d in data)
print('total warp-instructions', tot, 'stall samples', tots, 'sass lines', len(data))
top = sorted(data, key=lambda d: -d[1])[:topn]
thr = top[-1][1]
for d in data:
    if d[1] >= thr:
        print(f"{d[0] & 0xfffff:05x} {d[1]:>10.0f} {100*d[1]/tot:5.1f}% stall {100*d[2]/max(tots,1):5.1f}%  {d[3][:100]}")
