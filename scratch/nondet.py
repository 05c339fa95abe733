"""debug: run-to-run determinism of one ftrain (serial plan) -- outputs, loss, gradients"""
import os, sys
os.environ["MGCONV_LANES"] = os.environ.get("MGCONV_LANES", "1"); os.environ["MGCONV_AUTOTUNE"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multigrid-neural-architectures_b200"))
import torch
from mgconv import builders as B
torch.manual_seed(5)
net = B.load_net(sys.argv[2] if len(sys.argv) > 2 else "cifar/rnmg")
mn = "mnist" in (sys.argv[2] if len(sys.argv) > 2 else "")
pm = net.createModel(B.Opt(nLayer=1, nGPU=1, dataset="mnist-spt") if mn else B.Opt(nLayer=1, nGPU=1)); pm.precision = sys.argv[1]; pm.cuda()
if os.environ.get("BN_BIAS"):
    for m in pm.listModules():
        if m.typename == "nn.SpatialBatchNormalization":
            m.bias.fill_(float(os.environ["BN_BIAS"]))
params, grads = pm.getParameters()
crit = net.createCriterion()
g = torch.Generator(device="cpu").manual_seed(9)
if mn:
    x = torch.randn(8, 1, 64, 64, generator=g).cuda(); t = (torch.rand(8, 1, 64, 64, generator=g) < 0.1).float().cuda()
else:
    x = torch.randn(32, 3, 32, 32, generator=g).cuda(); t = torch.randint(1, 101, (32,), generator=g).cuda()
res = []
for trial in range(12):
    pm.zeroGradParameters()
    out, err = net.ftrain(x, t, pm, crit)
    torch.cuda.synchronize()
    res.append((out.clone(), float(err), grads.clone()))
o0, e0, g0 = res[0]
sizes = [(m.typename, w.numel()) for m in pm.listModules() for _, w, _ in m.own_parameters()]
for i, (o, e, gq) in enumerate(res):
    off, firstbad = 0, None
    for j, (tn, n) in enumerate(sizes):
        d = float((gq[off:off + n] - g0[off:off + n]).norm() / max(float(g0[off:off + n].norm()), 1e-20))
        if d > 1e-4 and n > 400 and firstbad is None: firstbad = (j, tn, n, round(d, 5))
        if d > 1e-4 and n > 400: lastbad = (j, tn, n, round(d, 5))
        off += n
    print(i, "out", float((o - o0).abs().max()), "loss", e - e0, "grad", float((gq - g0).norm() / g0.norm()), "first", firstbad, "last", lastbad if firstbad else None, flush=True)
