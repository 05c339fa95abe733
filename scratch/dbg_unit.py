import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multigrid-neural-architectures_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from mgconv import builders as B, ops as O
torch.manual_seed(1)
cin, cout, hs = [8,4,4],[8,4,4],[16,8,4]
pm = B.mgConv(list(cin), list(cout), [3]*3)
pm.precision = "fp32"; pm.needInputGrad = True
B.MSRinit(pm); B.BNinit(pm)
pm.cuda()
xs = [torch.randn(3,c,h,h).cuda() for c,h in zip(cin,hs)]
py = pm.forward(xs)
gi = pm.backward(None, [torch.randn_like(y) for y in py])
torch.cuda.synchronize()
E = pm._engine
for o in reversed(E.plan.ops):
    n = type(o).__name__
    info = [n, getattr(o,'name','')]
    for attr in ('dx','din','dcat'):
        if getattr(o, attr, None) is not None: info.append((attr, float(getattr(o,attr).float().abs().max())))
    if hasattr(o,'comb') and o.comb is not None: info.append(('comb', float(o.comb.buf.float().abs().max()), 'alias' if o.comb.alias else '', 'nsrc', len(o.comb.t.srcs)))
    if isinstance(o, O.ApplyOp):
        info.append(('dsums', o.dsums.abs().max().item() if o.dsums is not None else None))
        info.append(('mean', o.mean.abs().max().item(), 'invstd', o.invstd.abs().max().item(), 'coef', o.coef.abs().max().item()))
    if isinstance(o, O.ConvOp):
        info.append(('gw', o.mod.gradWeight.abs().max().item()))
    print(info)
