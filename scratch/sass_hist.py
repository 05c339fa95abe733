"""SASS instruction histogram of libmgconv.so per tensor-core kernel: python scratch/sass_hist.py [lib] > profiles/r2_sass_histogram.txt"""
import re, collections, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "multigrid-neural-architectures_b200/mgconv/libmgconv.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn = None; per = collections.defaultdict(collections.Counter); tot = collections.Counter()
for l in sass.splitlines():
    m = re.search(r'Function : (\S+)', l)
    if m:
        fn = m.group(1); continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
    if m and fn:
        op = m.group(1).split('.')[0]
        per[fn][op] += 1; tot[op] += 1
keys = ['UTCHMMA', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'LDTM', 'UTCBAR', 'SYNCS', 'LDGSTS', 'ELECT', 'UCGABAR_ARV']
def dem(n):
    d = subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
    d = d.replace('(anonymous namespace)::', '')
    m = re.search(r'(\w+(?:<[^>]*>)?)\(', d)
    return m.group(1) if m else d
print("SASS instruction histogram of the shipped libmgconv.so (cuobjdump -sass, sm_100a), round 2\n")
print("whole library: " + ", ".join(f"{k} {tot[k]}" for k in keys + ['HMMA', 'LDG', 'STG', 'REDG', 'ATOMG'] if tot[k]) + "  (HMMA = mma.sync: none)\n")
print(f"{'kernel':44s} " + " ".join(f"{k:>11s}" for k in keys))
for f, c in sorted(per.items(), key=lambda kv: (-kv[1]['UTCHMMA'], kv[0])):
    if not any(c[k] for k in keys[:6]): continue
    print(f"{dem(f)[:44]:44s} " + " ".join(f"{c[k]:11d}" for k in keys))
print("""
UTCHMMA = tcgen05.mma (kind::f16; the cta_group::2 form in the pair kernels), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
UTMALDG / UTMASTG = cp.async.bulk.tensor load / store through TMA tensor maps (halos, g tiles, stem patches, weight half-stages of
the pair kernels, stem output tiles), UBLKCP = cp.async.bulk 1-D (pre-swizzled weight stages), SYNCS = mbarrier operations,
UCGABAR_ARV = barrier.cluster (CTA pairs), LDGSTS = cp.async 16-byte gathers: only the generic per-tap kernels (1x1 convolutions, grids
wider than 64 columns, Linear head) still gather that way.""")
