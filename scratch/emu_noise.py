import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "multigrid-neural-architectures_b200"))
import numpy as np, torch, copy
from oracle import builders as OB
from util import rel_err, bf16_round, emulate_bf16_storage
torch.manual_seed(2); rng = np.random.default_rng(5)
N=16
om = OB.cifar_nmg(1).double()
om2 = copy.deepcopy(om); emulate_bf16_storage(om2)
x = bf16_round(rng.standard_normal((N,3,32,32))); t = rng.integers(1, 101, N)
for m in (om, om2):
    lp = m(torch.from_numpy(x)); torch.nn.functional.nll_loss(lp, torch.from_numpy(t-1)).backward()
import torch.nn as tnn
l1 = [m for m in om.modules() if isinstance(m,(tnn.Conv2d,tnn.Linear))]
l2 = [m for m in om2.modules() if isinstance(m,(tnn.Conv2d,tnn.Linear))]
for a,b in zip(l1,l2):
    print(tuple(a.weight.shape), "emulated-bf16 vs fp64 gw rel err %.4f" % rel_err(b.weight.grad.numpy(), a.weight.grad.numpy()))
