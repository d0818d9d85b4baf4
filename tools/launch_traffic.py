"""Per-kernel time and DRAM traffic of the LAST full ips() step in an ncu launch-list CSV
(metrics gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum).  Writes a JSON summary."""
import csv, json, re, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, mi, ii, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name'), hdr.index('ID'), hdr.index('Metric Unit')
d = {}
for r in rows[1:]:
    v = float(r[vi].replace(',', ''))
    u = r[ui]
    if u in ('Kbyte', 'KB'): v *= 1e3
    elif u in ('Mbyte', 'MB'): v *= 1e6
    elif u in ('Gbyte', 'GB'): v *= 1e9
    elif u == 'us': v *= 1e3
    elif u == 'ms': v *= 1e6
    d.setdefault(int(r[ii]), {'name': re.sub(r'[(].*', '', r[ki]).replace('void ', '').replace('<unnamed>::', '')})[r[mi]] = v
ids = sorted(d)
# a step ends with the final gather; take the launches between the last two gathers
ends = [k for k, i in enumerate(ids) if 'gather_rows' in d[i]['name']]
want = int(sys.argv[3]) if len(sys.argv) > 3 else 2            # staging launches of the step to report (2 = resident-input step)
s = e = None
for a, b in zip(ends[:-1], ends[1:]):
    if sum('stage_s2d' in d[i]['name'] for i in ids[a + 1:b + 1]) == want:
        s, e = a + 1, b + 1
assert s is not None, 'no step with %d staging launches' % want
agg = collections.OrderedDict()
for i in ids[s:e]:
    x = d[i]
    a = agg.setdefault(x['name'], {'launches': 0, 'us': 0.0, 'dram_read_MB': 0.0, 'dram_write_MB': 0.0})
    a['launches'] += 1
    a['us'] += x['gpu__time_duration.sum'] / 1e3
    a['dram_read_MB'] += x.get('dram__bytes_read.sum', 0) / 1e6
    a['dram_write_MB'] += x.get('dram__bytes_write.sum', 0) / 1e6
tot = sum(a['us'] for a in agg.values())
for k, a in agg.items():
    a['share'] = a['us'] / tot
    print(f"{a['us']:9.1f} us {100*a['share']:5.1f}%  n={a['launches']:3d}  rd {a['dram_read_MB']:8.1f} MB  wr {a['dram_write_MB']:8.1f} MB  {k[:70]}")
print(f'total {tot:.1f} us over {e - s} launches')
fam = [a for k, a in agg.items() if 'conv_' in k or 'stem_pool' in k]
out = {'source': sys.argv[1].split('/')[-1], 'step_us_serialised_cold': tot, 'kernels': agg,
       'conv_family': {'launches': sum(a['launches'] for a in fam), 'us': sum(a['us'] for a in fam),
                       'dram_bytes': sum(a['dram_read_MB'] + a['dram_write_MB'] for a in fam) * 1e6,
                       'share': sum(a['us'] for a in fam) / tot}}
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], 'w'), indent=1)
