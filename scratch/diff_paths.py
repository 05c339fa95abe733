"""debug: run one network forward + backward in bf16 and dump every plan buffer (conv y, dcat, gradWeight) to a file;
run twice with different MGCONV_* environments and compare with `diff_paths.py cmp a.pt b.pt`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multigrid-neural-architectures_b200"))
import numpy as np, torch

if sys.argv[1] == "cmp":
    a, b = torch.load(sys.argv[2]), torch.load(sys.argv[3])
    for k in a:
        x, y = a[k].float(), b[k].float()
        d = (x - y).abs().max().item()
        s = y.abs().max().item()
        flag = "  <<<<<<" if d > 0.05 * max(s, 1e-6) else ""
        print(f"{k[:70]:70s} maxdiff {d:.3e} scaleA {x.abs().max().item():.3e} scaleB {s:.3e}{flag}")
    sys.exit(0)

from mgconv import builders as B, ops
net = {"rnmg": B.cifar_rnmg, "nmg": B.cifar_nmg}[sys.argv[2]]
torch.manual_seed(2)
rng = np.random.default_rng(5)
pm = net.createModel(B.Opt(nGPU=1, nLayer=1))
pm.precision = "bf16"
pm.cuda()
x = torch.from_numpy(rng.standard_normal((16, 3, 32, 32)).astype(np.float32)).cuda()
t = torch.from_numpy(rng.integers(1, 101, 16)).cuda()
crit = net.createCriterion()
pm.zeroGradParameters()
out, err = net.ftrain(x, t, pm, crit)
torch.cuda.synchronize()
E = pm._engine
dump = {}
for i, o in enumerate(E.plan.ops):
    if isinstance(o, ops.ConvOp):
        tag = f"{i:03d} conv {o.name} H{o.H} Ccat{o.CcatP} Cout{o.Cout} k{o.k} segs{[(t_.Cp, m) for t_, m in o.segs]}"
        dump[tag + " y"] = o.y.buf.cpu()
        if o.dcat is not None:
            dump[tag + " dcat"] = o.dcat.cpu()
        dump[tag + " gW"] = o.mod.gradWeight.cpu().clone()
print("loss", float(err))
torch.save(dump, sys.argv[1])
