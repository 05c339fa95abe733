"""debug: does the lane schedule ever disagree with the serial plan?  (fp32 cifar/rnmg, one ftrain per trial)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multigrid-neural-architectures_b200"))
import torch
from mgconv import builders as B

def build(lanes, wlane):
    os.environ["MGCONV_LANES"] = lanes; os.environ["MGCONV_WGRAD_LANE"] = wlane; os.environ["MGCONV_AUTOTUNE"] = "0"
    torch.manual_seed(5)
    net = B.load_net("cifar/rnmg")
    pm = net.createModel(B.Opt(nLayer=1, nGPU=1)); pm.precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"; pm.cuda()
    params, grads = pm.getParameters()
    return net, pm, params, grads, net.createCriterion()

g = torch.Generator(device="cpu").manual_seed(9)
x = torch.randn(32, 3, 32, 32, generator=g).cuda(); t = torch.randint(1, 101, (32,), generator=g).cuda()
def run(net, pm, grads, crit):
    pm.zeroGradParameters()
    out, err = net.ftrain(x, t, pm, crit)
    torch.cuda.synchronize()
    return grads.clone()
net, pm, params, grads, crit = build("1", "1")
ref = run(net, pm, grads, crit)
ref2 = run(net, pm, grads, crit)
print("serial vs serial", float((ref - ref2).norm() / ref.norm()))
thr = float(os.environ.get("THR", "1e-5"))
for cfg in (("1", "0"), ("3", "1")):
    net, pm, params, grads, crit = build(*cfg)
    sizes = [(m.typename, w.numel()) for m in pm.listModules() for _, w, _ in m.own_parameters()]
    bad = 0
    for trial in range(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
        gq = run(net, pm, grads, crit)
        e = float((gq - ref).norm() / ref.norm())
        if e > thr:
            bad += 1
            off, first = 0, None
            worst = []
            for i, (tn, n) in enumerate(sizes):
                d = float((gq[off:off + n] - ref[off:off + n]).norm() / max(float(ref[off:off + n].norm()), 1e-20))
                if d > 1e-4: worst.append((i, tn, n, round(d, 5)))
                off += n
            print(cfg, "trial", trial, "err", e, "n_bad_params", len(worst), "first", worst[:3], "last", worst[-2:], flush=True)
    print(cfg, "bad", bad, "last err", e, flush=True)
