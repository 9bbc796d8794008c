"""ncu launch list (--metrics gpu__time_duration.sum, csv) -> per-kernel summary text.  usage: summarize_launches.py in.csv"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]
ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
tot = collections.OrderedDict()
n = 0
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r'\(.*', '', r[ki]).replace('void ', '').replace('<unnamed>::', '')
    v = float(r[vi].replace(',', '')) / 1e3
    n += 1
    tot.setdefault(name, [0, 0.0]); tot[name][0] += 1; tot[name][1] += v
T = sum(v for _, v in tot.values())
print(f'launches {n}, summed device time {T:.1f} us (ncu: cold cache, serialised -- compare SHARES, not absolutes)')
print(f'{"us":>10s} {"share":>6s} {"count":>5s}  kernel')
for k, (c, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f'{v:10.1f} {100 * v / T:5.1f}% {c:5d}  {k[:110]}')
