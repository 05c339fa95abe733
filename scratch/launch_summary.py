"""summarise an ncu `--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list by kernel"""
import csv, collections, re, sys, json
rows = list(csv.reader(open(sys.argv[1], errors='ignore')))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]
kn, mv, mn, mu, idc = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Name'), h.index('Metric Unit'), h.index('ID')
SC = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(',', '')) * SC.get(r[mu], 1.0)
    name = re.sub(r'\(.*', '', r[kn]).split('::')[-1]
    if r[mn] == 'gpu__time_duration.sum':
        agg[name][0] += 1
        agg[name][1] += v
    elif r[mn].startswith('dram__bytes'):
        agg[name][2] += v
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':60s} {'n':>5s} {'ms':>9s} {'share':>6s} {'DRAM GB':>9s} {'GB/s':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {v[0]:5d} {v[1] / 1e3:9.3f} {100 * v[1] / tot:5.1f}% {v[2] / 1e9:9.3f} {v[2] / 1e3 / max(v[1], 1e-9):7.0f}")
print(f"{'total':60s} {sum(v[0] for v in agg.values()):5d} {tot / 1e3:9.3f}        {sum(v[2] for v in agg.values()) / 1e9:9.3f}")
if len(sys.argv) > 2:   # conv traffic record for bench.py
    conv = {k: v for k, v in agg.items() if k.startswith('umma_') or k.startswith('wgrad_reduce')}
    json.dump({"source": sys.argv[1], "what": "dram__bytes_read.sum + dram__bytes_write.sum summed over the convolution launches (umma_* and wgrad_reduce) of one R-MG-34 step, B = 256",
               "conv_launches": sum(v[0] for v in conv.values()), "conv_ms": sum(v[1] for v in conv.values()) / 1e3,
               "dram_bytes_per_step": sum(v[2] for v in conv.values())}, open(sys.argv[2], 'w'), indent=1)
