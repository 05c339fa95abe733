"""kernel timeline of one laned R-MG-34 training step (CUPTI through torch.profiler): busy time, idle gaps and concurrency
python scratch/step_trace.py [out.json]"""
import os, sys, json, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multigrid-neural-architectures_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from mgconv import builders as B

net = B.load_net("ilsvrc/rnmg")
torch.manual_seed(2)
model = net.createModel(B.Opt(nGPU=1, depth=34)); model.precision = "bf16"; model.cuda()
crit = net.createCriterion()
params, grads = model.getParameters()
st = dict(learningRate=0.05, momentum=0.9, weightDecay=1e-4, dampening=0.0)
x = torch.randn(256, 3, 224, 224, device="cuda"); t = torch.randint(1, 1001, (256,), device="cuda")
def step():
    model.zeroGradParameters()
    def feval(_p):
        out, err = net.ftrain(x, t, model, crit)
        return err, grads
    net.btrain(params, feval, st)
for _ in range(4): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2): step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
ks = sorted([(e.time_range.start, e.time_range.end, e.name) for e in ev if "memcpy" not in e.name.lower() and "memset" not in e.name.lower()])
# second step only: split at the largest gap... take the last half by count
half = len(ks) // 2
ks = ks[half:]
t0, t1 = ks[0][0], max(k[1] for k in ks)
print(f"kernels {len(ks)}  span {(t1 - t0) / 1e3:.2f} ms")
# union busy time and concurrency histogram
pts = []
for s, e, _ in ks: pts += [(s, 1), (e, -1)]
pts.sort()
cur, last, hist = 0, t0, collections.Counter()
for tt, d in pts:
    hist[cur] += tt - last; last = tt; cur += d
tot = sum(hist.values())
print("concurrency (kernels in flight): " + ", ".join(f"{k}: {v / 1e3:.2f} ms ({100 * v / tot:.0f} %)" for k, v in sorted(hist.items())))
import re
def fam(n):
    m = re.search(r"(\w+)(<[^>]*>)?\(", n.replace("(anonymous namespace)::", "").replace("<unnamed>::", ""))
    return (m.group(1) + (m.group(2) or "")) if m else n[:40]
# total time by family
totf = collections.Counter()
for s_, e_, n_ in ks: totf[fam(n_)] += e_ - s_
# time by kernel family while it is the ONLY kernel running
alone = collections.Counter(); cur = []
ev2 = []
for s, e, n in ks: ev2 += [(s, 0, n), (e, 1, n)]
ev2.sort()
active, last = [], t0
for tt, kind, n in ev2:
    if len(active) == 1: alone[fam(active[0])] += tt - last
    last = tt
    if kind == 0: active.append(n)
    else: active.remove(n)
print("time a kernel runs ALONE, by family:")
for n, v in alone.most_common(16): print(f"   {n:42s} alone {v / 1e3:7.2f} ms   of {totf[n] / 1e3:7.2f} ms total")
if len(sys.argv) > 1: json.dump(ks, open(sys.argv[1], "w"))
