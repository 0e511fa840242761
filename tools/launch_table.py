"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum,launch__grid_size,sm__cycles_active.avg,sm__cycles_elapsed.avg
--csv) by kernel: launches, device time, device time weighted by the fraction of cycles the SMs were active (what a kernel costs
when other streams fill the gaps). usage: launch_table.py file.csv [rows] [last N launches only]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, mi, ii = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name'), hdr.index('ID')
data = collections.OrderedDict()
for r in rows[1:]:
    d = data.setdefault(r[ii], {'name': r[ki]})
    d[r[mi]] = float(r[vi].replace(',', ''))
if len(sys.argv) > 3:
    keep = list(data.keys())[-int(sys.argv[3]):]
    data = collections.OrderedDict((k, data[k]) for k in keep)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
tot = totw = 0.0
for d in data.values():
    t = d.get('gpu__time_duration.sum', 0) / 1000.0
    act, el = d.get('sm__cycles_active.avg', 0), d.get('sm__cycles_elapsed.avg', 1)
    w = t * act / el if el else t
    name = d['name'].split('(')[0][:48]
    a = agg[name]
    a[0] += 1; a[1] += t; a[2] = max(a[2], d.get('launch__grid_size', 0)); a[3] += w
    tot += t; totw += w
print('| us | SM-active-weighted us | launches | max grid | kernel |\n|---:|---:|---:|---:|---|')
print('| %.1f | %.1f | %d | | **total** |' % (tot, totw, len(data)))
for n, a in sorted(agg.items(), key=lambda x: -x[1][3])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print('| %.1f | %.1f | %d | %d | `%s` |' % (a[1], a[3], a[0], a[2], n))
