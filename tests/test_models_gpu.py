"""Whole-unit and whole-network parity on the GPU: the product's builder API (mgconv.builders,
the mirror of models/*.lua) against the oracle's restatement of the same Lua builders
(oracle/builders.py on PyTorch-CPU, float64), same weights, same inputs.

fp32 mode: rel 1e-4; bf16 mode: rel 2e-2 per residual unit (BASELINE.json north_star).  Whole
networks in bf16 stack 15-70 roundings of activations and gradients, so their bound is the unit
tolerance scaled by sqrt(depth) -- stated per case below.
"""
import math
import numpy as np
import pytest
import torch

from oracle import builders as OB
from mgconv import builders as B, nn
from util import copy_params_from_oracle, rel_err, bf16_round

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-4, "bf16": 2e-2}


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _param_grads(olist, plist):
    og = np.concatenate([np.concatenate([m.weight.grad.numpy().ravel(), m.bias.grad.numpy().ravel()]) for m in olist])
    pg = np.concatenate([np.concatenate([m.gradWeight.cpu().numpy().ravel(), m.gradBias.cpu().numpy().ravel()]) for m in plist])
    return og, pg


def _is_affine(m):
    return isinstance(m, torch.nn.BatchNorm2d)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("cin,cout,hs", [([8, 4, 4], [8, 4, 4], [16, 8, 4]),       # identity shortcuts
                                           ([8, 4, 4], [16, 8, 4], [8, 4, 2]),       # zero-padded shortcuts
                                           ([12, 6], [12, 6], [7, 4]),               # odd finer grid: ceil-mode pooling
                                           ([16], [24], [5])])                       # single grid = plain residual pair
def test_residual_mg_unit(precision, cin, cout, hs):
    """models/ilsvrc/rnmg.lua:91-159"""
    torch.manual_seed(1)
    rng = np.random.default_rng(3)
    N = 3
    om = OB.res_mgConv(list(cin), list(cout), [3] * len(cin)).double()
    for m in om.modules():
        if _is_affine(m):
            m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.1)
    pm = B.mgConv(list(cin), list(cout), [3] * len(cin))
    pm.precision = precision
    pm.needInputGrad = True
    olist, plist = copy_params_from_oracle(om, pm)
    pm.cuda()
    plist = [m for m in pm.listModules() if m.own_parameters()]
    xs = [bf16_round(rng.standard_normal((N, c, h, h))) for c, h in zip(cin, hs)]
    oxs = [_t(x).requires_grad_() for x in xs]
    oy = om(oxs)
    gos = [bf16_round(rng.standard_normal(tuple(y.shape))) for y in oy]
    torch.autograd.backward(oy, [_t(g) for g in gos])

    py = pm.forward([_t(x).float().cuda() for x in xs])
    gi = pm.backward(None, [_t(g).float().cuda() for g in gos])
    torch.cuda.synchronize()
    tol = TOL[precision]
    gi = gi if isinstance(gi, list) else [gi]
    for i in range(len(cin)):
        assert rel_err(py[i].cpu().numpy(), oy[i].detach().numpy()) <= tol, ("out", i)
        assert rel_err(gi[i].cpu().numpy(), oxs[i].grad.numpy()) <= 2 * tol, ("gradInput", i)
    for o, p in zip(olist, plist):
        assert rel_err(p.gradWeight.cpu().numpy(), o.weight.grad.numpy()) <= 2 * tol, (p.typename, "gradWeight")
        if p.typename != "cudnn.SpatialConvolution":  # conv bias gradient under BN is ~0 (pure rounding noise)
            assert rel_err(p.gradBias.cpu().numpy(), o.bias.grad.numpy()) <= 2 * tol, (p.typename, "gradBias")
        else:
            assert np.abs(p.gradBias.cpu().numpy()).max() <= 50 * tol * max(1.0, np.abs(o.weight.grad.numpy()).max())


NETS = [
    # name, oracle ctor, product NET, opt, input shape, nClass, depth for the bf16 bound
    ("MG-6 cifar/nmg nLayer=1", lambda: OB.cifar_nmg(1), B.cifar_nmg, dict(nLayer=1), (4, 3, 32, 32), 100, 6),
    ("R-NMG-12 cifar/rnmg nLayer=1", lambda: OB.cifar_rnmg(1), B.cifar_rnmg, dict(nLayer=1), (4, 3, 32, 32), 100, 12),
    ("PR-NMG-16 cifar/prnmg nLayer=1 narrow", lambda: OB.cifar_prnmg(1, blocks=OB.CIFAR_RNMG_NARROW), B.cifar_prnmg,
     dict(nLayer=1, blocks=B.CIFAR_RNMG_BLOCKS), (4, 3, 32, 32), 100, 16),
    ("R-MG-10 ilsvrc/rnmg reduced 64x64",
     lambda: OB.ilsvrc_rnmg(18, 10, [16, 8, 8], [([16, 8, 8], [3, 3, 3], False), ([32, 16, 8], [3, 3, 3], True), ([32, 16], [3, 3], True), ([64], [3], False)], [1, 1, 1, 1], 2),
     B.ilsvrc_rnmg, dict(nClass=10, inputBlock=[16, 8, 8], cfg=[1, 1, 1, 1], avg=2,
                         blocks=[([16, 8, 8], [3, 3, 3], False), ([32, 16, 8], [3, 3, 3], True), ([32, 16], [3, 3], True), ([64], [3], False)]),
     (3, 3, 64, 64), 10, 10),
]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("case", range(len(NETS)), ids=[n[0] for n in NETS])
def test_network_forward_backward(precision, case):
    name, octor, net, opt, shape, nclass, depth = NETS[case]
    torch.manual_seed(2)
    rng = np.random.default_rng(5)
    om = octor().double()
    pm = net.createModel(B.Opt(nGPU=1, **opt))
    pm.precision = precision
    olist, plist = copy_params_from_oracle(om, pm)
    pm.cuda()
    plist = [m for m in pm.listModules() if m.own_parameters()]
    x = bf16_round(rng.standard_normal(shape))
    t = rng.integers(1, nclass + 1, shape[0])
    olp = om(_t(x))
    oloss = torch.nn.functional.nll_loss(olp, _t(t - 1))
    oloss.backward()

    crit = net.createCriterion()
    xd, td = _t(x).float().cuda(), _t(t).cuda()
    outputs, err = net.ftrain(xd, td, pm, crit)
    torch.cuda.synchronize()
    tol = TOL[precision] * (1 if precision == "fp32" else math.sqrt(depth))
    assert rel_err(outputs.cpu().numpy(), olp.detach().numpy()) <= tol, "log-probabilities"
    assert abs(float(err) - oloss.item()) <= tol * max(1.0, abs(oloss.item())), "loss"
    og, pg = [], []
    for o, p in zip(olist, plist):
        og.append(o.weight.grad.numpy().ravel()); pg.append(p.gradWeight.cpu().numpy().ravel())
        if p.typename != "cudnn.SpatialConvolution":
            og.append(o.bias.grad.numpy().ravel()); pg.append(p.gradBias.cpu().numpy().ravel())
    og, pg = np.concatenate(og), np.concatenate(pg)
    assert rel_err(pg, og) <= 4 * tol, "parameter gradients"
    # running statistics were updated like nn.SpatialBatchNormalization does (momentum 0.1, unbiased var)
    obn = [m for m in olist if _is_affine(m)][0]
    pbn = [m for m in plist if m.typename == "nn.SpatialBatchNormalization"][0]
    assert rel_err(pbn.running_var.cpu().numpy(), obn.running_var.numpy()) <= tol

    # evaluation mode: running statistics, no backward
    om.eval(); pm.evaluate()
    with torch.no_grad():
        olp2 = om(_t(x))
    assert rel_err(pm.forward(xd).cpu().numpy(), olp2.numpy()) <= tol, "evaluate()"


def test_mnist_prnmg_dense_prediction():
    """models/mnist-cluttered/prnmg.mnist.lua: shrinking pyramid (isDrop), 1x1 conv+BN shortcuts,
    last unit without ReLU (isOut), Sigmoid + BCECriterion"""
    torch.manual_seed(4)
    rng = np.random.default_rng(6)
    om = OB.mnist_prnmg(1, 1).double()
    pm = B.mnist_prnmg.createModel(B.Opt(nLayer=1, nGPU=1, dataset="mnist-spt"))
    pm.precision = "fp32"
    olist, plist = copy_params_from_oracle(om, pm)
    pm.cuda()
    plist = [m for m in pm.listModules() if m.own_parameters()]
    x = bf16_round(rng.standard_normal((2, 1, 64, 64)))
    t = (rng.random((2, 1, 64, 64)) < 0.1).astype(np.float64)
    op = om(_t(x))
    oloss = torch.nn.functional.binary_cross_entropy(op, _t(t))
    oloss.backward()
    crit = B.mnist_prnmg.createCriterion()
    out, err = B.mnist_prnmg.ftrain(_t(x).float().cuda(), _t(t).float().cuda(), pm, crit)
    torch.cuda.synchronize()
    assert rel_err(out.cpu().numpy(), op.detach().numpy()) <= 1e-4
    assert abs(float(err) - oloss.item()) <= 1e-4
    og = np.concatenate([o.weight.grad.numpy().ravel() for o in olist])
    pg = np.concatenate([p.gradWeight.cpu().numpy().ravel() for p in plist])
    assert rel_err(pg, og) <= 1e-3


def test_train_steps_follow_the_oracle():
    """three ftrain + btrain steps (optim.sgd momentum .9, wd 5e-4; models/basic_model.lua:56-66)"""
    torch.manual_seed(3)
    rng = np.random.default_rng(8)
    om = OB.cifar_nmg(1).double()
    pm = B.cifar_nmg.createModel(B.Opt(nLayer=1, nGPU=1))
    pm.precision = "fp32"
    copy_params_from_oracle(om, pm)
    pm.cuda()
    params, grads = pm.getParameters()
    crit = B.cifar_nmg.createCriterion()
    opt = torch.optim.SGD(om.parameters(), lr=0.05, momentum=0.9, weight_decay=5e-4)
    state = dict(learningRate=0.05, momentum=0.9, weightDecay=5e-4, dampening=0.0)
    for step in range(3):
        x = bf16_round(rng.standard_normal((4, 3, 32, 32)))
        t = rng.integers(1, 101, 4)
        opt.zero_grad()
        loss = torch.nn.functional.nll_loss(om(_t(x)), _t(t - 1))
        loss.backward(); opt.step()
        pm.zeroGradParameters()
        xd, td = _t(x).float().cuda(), _t(t).cuda()

        def feval(p):
            outputs, err = B.cifar_nmg.ftrain(xd, td, pm, crit)
            return err, grads
        _, fx = B.cifar_nmg.btrain(params, feval, state)
        assert abs(float(fx[0]) - loss.item()) <= 1e-3 * max(1, abs(loss.item())), step
    ow = np.concatenate([p.detach().numpy().ravel() for p in om.parameters()])
    assert rel_err(params.cpu().numpy(), ow) <= 1e-4


def test_no_cpu_fallback():
    from mgconv import ffi
    m = B.cifar_nmg.createModel(B.Opt(nLayer=1))
    with pytest.raises(ffi.MGError):
        m.forward(torch.zeros(1, 3, 32, 32))  # CPU tensor
