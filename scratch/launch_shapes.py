"""per (kernel, grid, block) totals of an ncu launch list: python scratch/launch_shapes.py <csv> [name-prefix ...]"""
import csv, re, collections, sys
rows = list(csv.reader(open(sys.argv[1], errors='ignore')))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; kn, mv, mn, mu, gs, bs = (h.index(x) for x in ('Kernel Name', 'Metric Value', 'Metric Name', 'Metric Unit', 'Grid Size', 'Block Size'))
SC = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
pre = sys.argv[2:]
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    name = re.sub(r'\(.*', '', r[kn]).split('::')[-1]
    if pre and not any(name.startswith(p) for p in pre): continue
    v = float(r[mv].replace(',', '')) * SC.get(r[mu], 1.0)
    k = (name, r[gs], r[bs])
    if r[mn] == 'gpu__time_duration.sum': agg[k][0] += 1; agg[k][1] += v
    elif r[mn].startswith('dram__bytes'): agg[k][2] += v
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[0][:38]:38s} {k[1]:16s} {k[2]:13s} n={v[0]:3d} tot={v[1]:8.1f}us avg={v[1]/v[0]:7.1f}us  {v[2]/v[0]/1e6:8.1f} MB/launch {v[2]/1e3/max(v[1],1e-9):6.0f} GB/s")
