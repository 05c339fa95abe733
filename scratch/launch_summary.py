"""summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1], errors='ignore')))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]
kn, mv, mn, mu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Name'), h.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv or r[mn] != 'gpu__time_duration.sum':
        continue
    v = float(r[mv].replace(',', ''))
    v = v / 1e3 if r[mu] == 'ns' else (v * 1e3 if r[mu] == 'ms' else v)
    name = re.sub(r'\(.*', '', r[kn]).split('::')[-1]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':60s} {'n':>5s} {'ms':>9s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {v[0]:5d} {v[1] / 1e3:9.3f} {100 * v[1] / tot:5.1f}%")
print(f"{'total':60s} {sum(v[0] for v in agg.values()):5d} {tot / 1e3:9.3f}")
