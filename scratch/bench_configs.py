"""secondary BASELINE.json configs (2, 3, 5): train-step throughput of the CIFAR / MNIST nets on 1 GPU"""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multigrid-neural-architectures_b200"))
import torch
from mgconv import builders as B

CONFIGS = {
    "R-MG-22 cifar/rnmg nLayer=2 B=128": ("cifar/rnmg", dict(nLayer=2), (128, 3, 32, 32), "nll", 100),
    "PR-NMG-30 cifar/prnmg nLayer=2 B=256": ("cifar/prnmg", dict(nLayer=2), (256, 3, 32, 32), "nll", 100),
    "MG-6 cifar/nmg nLayer=1 B=64": ("cifar/nmg", dict(nLayer=1), (64, 3, 32, 32), "nll", 100),
    "PR-NMG mnist-cluttered/prnmg.mnist B=128": ("mnist-cluttered/prnmg.mnist", dict(nLayer=1, dataset="mnist-spt"), (128, 1, 64, 64), "bce", 1),
    "U-MG mnist-cluttered/unmg B=128": ("mnist-cluttered/unmg", dict(dataset="mnist-seg"), (128, 1, 64, 64), "bce", 10),
}
use_graph = "--graph" in sys.argv
for name, (nt, opt, shape, loss, ncls) in CONFIGS.items():
    torch.manual_seed(2)
    net = B.load_net(nt)
    model = net.createModel(B.Opt(nGPU=1, **opt)); model.precision = "bf16"; model.cuda()
    crit = net.createCriterion()
    params, grads = model.getParameters()
    st = dict(learningRate=0.05, momentum=0.9, weightDecay=5e-4, dampening=0.0)
    x = torch.randn(*shape, device="cuda")
    t = torch.randint(1, ncls + 1, (shape[0],), device="cuda") if loss == "nll" else (torch.rand(shape[0], ncls, shape[2], shape[3], device="cuda") < 0.1).float()
    def step():
        model.zeroGradParameters()
        def feval(_p):
            out, err = net.ftrain(x, t, model, crit)
            return err, grads
        net.btrain(params, feval, st)
    for _ in range(3): step()
    torch.cuda.synchronize()
    run = step
    if use_graph:
        from mgconv.graph import GraphedStep
        run = GraphedStep(step)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    eng = model._engine
    print(json.dumps({"config": name, "ms_per_step": round(ms, 3), "images_per_s": round(shape[0] / ms * 1e3, 1), "graph": use_graph,
                      "launches_per_step": None}), flush=True)
