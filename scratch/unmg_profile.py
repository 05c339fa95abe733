"""one training step of mnist-cluttered/unmg (B = 128) under ncu / for timing"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multigrid-neural-architectures_b200"))
import torch
from mgconv import builders as B
torch.manual_seed(2)
net = B.load_net("mnist-cluttered/unmg")
model = net.createModel(B.Opt(nGPU=1, dataset="mnist-seg")); model.precision = "bf16"; model.cuda()
crit = net.createCriterion()
params, grads = model.getParameters()
st = dict(learningRate=0.05, momentum=0.9, weightDecay=5e-4, dampening=0.0)
x = torch.randn(128, 1, 64, 64, device="cuda")
t = (torch.rand(128, 10, 64, 64, device="cuda") < 0.1).float()
def step():
    model.zeroGradParameters()
    def feval(_p):
        out, err = net.ftrain(x, t, model, crit)
        return err, grads
    net.btrain(params, feval, st)
for _ in range(3): step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
