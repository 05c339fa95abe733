import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multigrid-neural-architectures_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import builders as OB
from mgconv import builders as B
from util import copy_params_from_oracle, rel_err, bf16_round, emulate_bf16_storage
precision = sys.argv[1]; emu = int(sys.argv[2]); N = int(sys.argv[3])
torch.manual_seed(2); rng = np.random.default_rng(5)
om = OB.cifar_nmg(1).double()
pm = B.cifar_nmg.createModel(B.Opt(nGPU=1, nLayer=1)); pm.precision = precision
olist, plist = copy_params_from_oracle(om, pm)
if emu: emulate_bf16_storage(om)
pm.cuda(); plist = [m for m in pm.listModules() if m.own_parameters()]
x = bf16_round(rng.standard_normal((N,3,32,32))); t = rng.integers(1, 101, N)
olp = om(torch.from_numpy(x)); oloss = torch.nn.functional.nll_loss(olp, torch.from_numpy(t-1)); oloss.backward()
crit = B.cifar_nmg.createCriterion()
out, err = B.cifar_nmg.ftrain(torch.from_numpy(x).float().cuda(), torch.from_numpy(t).cuda(), pm, crit)
torch.cuda.synchronize()
print("out", rel_err(out.cpu().numpy(), olp.detach().numpy()))
for i,(o,p) in enumerate(zip(olist, plist)):
    print(i, p.typename, tuple(p.weight.shape), "gw %.4f" % rel_err(p.gradWeight.cpu().numpy(), o.weight.grad.numpy()), "gb %.4f" % rel_err(p.gradBias.cpu().numpy(), o.bias.grad.numpy()))
