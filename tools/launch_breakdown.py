"""Per-kernel durations of one encoder chunk from an ncu launch-list CSV (gpu__time_duration.sum, sm__cycles_elapsed.max)."""
import csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, mi, ii = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name'), hdr.index('ID')
d = {}
for r in rows[1:]:
    d.setdefault(r[ii], {'name': r[ki]})[r[mi]] = float(r[vi].replace(',', ''))
ids = sorted(d, key=int)
starts = [k for k, i in enumerate(ids) if 'stage_s2d' in d[i]['name'] or 'stage_padded' in d[i]['name']]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1
s, e = starts[which], starts[which + 1]
tot = 0
for i in ids[s:e]:
    x = d[i]
    t = x['gpu__time_duration.sum'] / 1000
    c = x.get('sm__cycles_elapsed.max', 0)
    print(f"{t:8.1f} us {c / t / 1000:6.2f} GHz  {re.sub(r'[(].*', '', x['name'])[:70]}")
    tot += t
print(f'total {tot:.1f} us')
