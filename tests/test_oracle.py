"""Oracle self-checks (CPU): the numpy restatement (oracle/nn_ops.py) against PyTorch-CPU
op by op, and the builder restatement (oracle/builders.py) against the reference's
published structural numbers (README.md:85-92,109; models/ilsvrc/rnmg.lua:241,251-254)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import nn_ops as O
from oracle import builders as B

rng = np.random.default_rng(2)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a))


@pytest.mark.parametrize("H,W", [(8, 8), (7, 7), (5, 9), (1, 1), (2, 3)])
def test_maxpool_ceil_indices_bit_exact(H, W):
    x = rng.standard_normal((2, 3, H, W))
    x[0, 0, : min(2, H), : min(2, W)] = 1.5  # ties: first max in row-major scan wins
    y, idx = O.maxpool_forward(x)
    ty, tidx = F.max_pool2d(T(x), 2, 2, 0, ceil_mode=True, return_indices=True)
    assert np.array_equal(y, ty.numpy())
    assert np.array_equal(idx, tidx.numpy())
    g = rng.standard_normal(y.shape)
    xt = T(x).requires_grad_()
    F.max_pool2d(xt, 2, 2, 0, ceil_mode=True).backward(T(g))
    assert np.allclose(O.maxpool_backward(g, idx, x.shape), xt.grad.numpy())


def test_stem_maxpool_3x3s2p1():
    x = rng.standard_normal((2, 2, 12, 12))
    y, idx = O.maxpool_forward(x, 3, 2, 1, ceil_mode=False)
    ty, tidx = F.max_pool2d(T(x), 3, 2, 1, return_indices=True)
    assert np.array_equal(y, ty.numpy()) and np.array_equal(idx, tidx.numpy())
    g = rng.standard_normal(y.shape)
    xt = T(x).requires_grad_()
    F.max_pool2d(xt, 3, 2, 1).backward(T(g))
    assert np.allclose(O.maxpool_backward(g, idx, x.shape), xt.grad.numpy())


def test_upsample_and_avgpool():
    x = rng.standard_normal((2, 3, 4, 5))
    assert np.array_equal(O.upsample_forward(x), F.interpolate(T(x), scale_factor=2, mode="nearest").numpy())
    g = rng.standard_normal((2, 3, 8, 10))
    xt = T(x).requires_grad_()
    F.interpolate(xt, scale_factor=2, mode="nearest").backward(T(g))
    assert np.allclose(O.upsample_backward(g), xt.grad.numpy())
    x = rng.standard_normal((2, 3, 9, 8))
    assert np.allclose(O.avgpool_forward(x, 2), F.avg_pool2d(T(x), 2, 2).numpy())
    assert np.allclose(O.avgpool_forward(x, 7, 1), F.avg_pool2d(T(x), 7, 1).numpy())


@pytest.mark.parametrize("k,s,p", [(3, 1, 1), (1, 1, 0), (7, 2, 3)])
def test_conv_fwd_bwd(k, s, p):
    x = rng.standard_normal((2, 5, 9, 10))
    w = rng.standard_normal((4, 5, k, k))
    b = rng.standard_normal(4)
    y = O.conv_forward(x, w, b, s, p)
    xt, wt, bt = T(x).requires_grad_(), T(w).requires_grad_(), T(b).requires_grad_()
    ty = F.conv2d(xt, wt, bt, s, p)
    assert np.allclose(y, ty.detach().numpy())
    g = rng.standard_normal(y.shape)
    ty.backward(T(g))
    gx, gw, gb = O.conv_backward(x, w, g, s, p)
    assert np.allclose(gx, xt.grad.numpy()) and np.allclose(gw, wt.grad.numpy()) and np.allclose(gb, bt.grad.numpy())


def test_upconv():
    x = rng.standard_normal((2, 3, 4, 4)); w = rng.standard_normal((3, 5, 2, 2)); b = rng.standard_normal(5)
    xt, wt = T(x).requires_grad_(), T(w).requires_grad_()
    ty = F.conv_transpose2d(xt, wt, T(b), 2)
    assert np.allclose(O.upconv2x2_forward(x, w, b), ty.detach().numpy())
    g = rng.standard_normal(ty.shape); ty.backward(T(g))
    gx, gw, gb = O.upconv2x2_backward(x, w, g)
    assert np.allclose(gx, xt.grad.numpy()) and np.allclose(gw, wt.grad.numpy())


@pytest.mark.parametrize("eps", [1e-5, 1e-3])
def test_batchnorm(eps):
    x = rng.standard_normal((4, 3, 5, 6)) * 2 + 1
    gamma, beta = rng.random(3), rng.standard_normal(3)
    rm, rv = np.zeros(3), np.ones(3)
    y, mean, invstd = O.bn_forward_train(x, gamma, beta, eps, rm, rv)
    xt, gt, bt = T(x).requires_grad_(), T(gamma).requires_grad_(), T(beta).requires_grad_()
    trm, trv = torch.zeros(3, dtype=torch.float64), torch.ones(3, dtype=torch.float64)
    ty = F.batch_norm(xt, trm, trv, gt, bt, True, 0.1, eps)
    assert np.allclose(y, ty.detach().numpy())
    assert np.allclose(rm, trm.numpy()) and np.allclose(rv, trv.numpy())  # unbiased var in running stats
    g = rng.standard_normal(y.shape); ty.backward(T(g))
    gx, dg, db = O.bn_backward_train(x, g, gamma, mean, invstd)
    assert np.allclose(gx, xt.grad.numpy()) and np.allclose(dg, gt.grad.numpy()) and np.allclose(db, bt.grad.numpy())
    assert np.allclose(O.bn_forward_eval(x, gamma, beta, rm, rv, eps),
                       F.batch_norm(T(x), trm, trv, T(gamma), T(beta), False, 0.1, eps).numpy())


def _rand_params(cins, couts, k=3):
    return [dict(w=rng.standard_normal((co, ci, k, k)) * 0.2, b=rng.standard_normal(co) * 0.1,
                 gamma=rng.random(co) + 0.5, beta=rng.standard_normal(co) * 0.1) for ci, co in zip(cins, couts)]


def _torch_stage(xs, params, eps, relu):
    n = len(xs)
    ys = []
    for i in range(n):
        parts = []
        if i > 0: parts.append(F.max_pool2d(xs[i - 1], 2, 2, 0, ceil_mode=True))
        parts.append(xs[i])
        if i + 1 < n: parts.append(F.interpolate(xs[i + 1], scale_factor=2, mode="nearest"))
        p = params[i]
        c = F.conv2d(torch.cat(parts, 1), p["w"], p["b"], 1, 1)
        c = F.batch_norm(c, None, None, p["gamma"], p["beta"], True, 0.1, eps)
        ys.append(F.relu(c) if relu else c)
    return ys


def test_residual_mg_unit_fwd_bwd_vs_autograd():
    """models/ilsvrc/rnmg.lua:91-159 incl. zero-padded shortcut and ConcatTable grad summation"""
    cin, cout, hs = [4, 2, 2], [6, 4, 2], [8, 4, 2]
    xs = [rng.standard_normal((2, c, h, h)) for c, h in zip(cin, hs)]
    cat1 = [cin[0] + cin[1], sum(cin), cin[1] + cin[2]]
    cat2 = [cout[0] + cout[1], sum(cout), cout[1] + cout[2]]
    p1, p2 = _rand_params(cat1, cout), _rand_params(cat2, cout)
    outs, ctx = O.mg_resunit_forward(xs, p1, p2)
    gos = [rng.standard_normal(o.shape) for o in outs]
    gx, gp1, gp2 = O.mg_resunit_backward(gos, p1, p2, ctx)

    txs = [T(x).requires_grad_() for x in xs]
    tp1 = [{k: T(v).requires_grad_() for k, v in p.items()} for p in p1]
    tp2 = [{k: T(v).requires_grad_() for k, v in p.items()} for p in p2]
    h = _torch_stage(txs, tp1, 1e-5, True)
    z = _torch_stage(h, tp2, 1e-5, False)
    touts = [F.relu(z[i] + F.pad(txs[i], (0, 0, 0, 0, 0, cout[i] - cin[i]))) for i in range(3)]
    torch.autograd.backward(touts, [T(g) for g in gos])
    for i in range(3):
        assert np.allclose(outs[i], touts[i].detach().numpy())
        assert np.allclose(gx[i], txs[i].grad.numpy())
        for k in ("w", "b", "gamma", "beta"):
            assert np.allclose(gp1[i][k], tp1[i][k].grad.numpy(), atol=1e-9), (i, k)
            assert np.allclose(gp2[i][k], tp2[i][k].grad.numpy(), atol=1e-9), (i, k)


def test_head_and_criteria():
    x = rng.standard_normal((5, 7)); t = rng.integers(0, 7, 5)
    lp = O.logsoftmax_forward(x)
    xt = T(x).requires_grad_()
    loss = F.nll_loss(F.log_softmax(xt, 1), T(t)); loss.backward()
    assert np.allclose(O.nll_forward(lp, t), loss.item())
    assert np.allclose(O.logsoftmax_backward(lp, O.nll_backward(lp, t)), xt.grad.numpy())
    p = rng.random((3, 4)); tt = (rng.random((3, 4)) > 0.5).astype(float)
    pt = T(p).requires_grad_(); l = F.binary_cross_entropy(pt, T(tt)); l.backward()
    assert np.allclose(O.bce_forward(p, tt), l.item(), atol=1e-9)
    assert np.allclose(O.bce_backward(p, tt), pt.grad.numpy(), atol=1e-8)


def test_sgd_matches_optim_sgd_semantics():
    w = rng.standard_normal(10); st = {}
    tw = T(w.copy()).requires_grad_()
    opt = torch.optim.SGD([tw], lr=0.1, momentum=0.9, weight_decay=1e-4, dampening=0)
    for _ in range(3):
        g = rng.standard_normal(10)
        w = O.sgd_step(w, g, st, 0.1, 0.9, 1e-4)
        tw.grad = T(g.copy()); opt.step()
    assert np.allclose(w, tw.detach().numpy())


# ---- structural known answers published by the reference -----------------------------
KNOWN = [  # (builder, input, params, MACs)  README.md:85-92,109 + SURVEY.md §6 re-derivation
    (lambda: B.cifar_nmg(1), (1, 3, 32, 32), 3_361_980, 51_707_520),
    (lambda: B.cifar_rnmg(2), (1, 3, 32, 32), 17_524_920, 397_059_840),
    (lambda: B.cifar_rnmg(2, blocks=B.CIFAR_WIDE), (1, 3, 32, 32), 44_789_316, 1_015_262_208),  # README R-MG-22: 44.79M
    (lambda: B.cifar_prnmg(2), (1, 3, 32, 32), 44_882_244, 1_031_777_280),
    (lambda: B.ilsvrc_rnmg(34), (1, 3, 224, 224), 32_899_176, 5_759_765_760),  # README: 32.9M, 5.76G
]


@pytest.mark.parametrize("case", range(len(KNOWN)))
def test_builder_known_answers(case):
    mk, shape, params, macs = KNOWN[case]
    torch.manual_seed(0)
    m = mk()
    assert B.count_params(m) == params
    assert B.count_conv_macs(m, torch.randn(*shape)) == macs


def test_ilsvrc_stage_shapes():
    """(224,112,56)->(56,28,14) ; (56,28,14)->(28,14,7) ; (28,14,7)->(14,7) ; (14,7)->(7)
    models/ilsvrc/rnmg.lua:241,251-254"""
    torch.manual_seed(0)
    m = B.ilsvrc_rnmg(18)
    x = torch.randn(1, 3, 224, 224)
    shapes = []
    t = x
    for mod in list(m)[:-1]:
        t = mod(t)
        shapes.append([(e.shape[1], e.shape[2]) for e in t])
    assert shapes[0] == [(64, 56), (32, 28), (16, 14)]
    assert [(128, 14), (96, 7)] in shapes and [(384, 7)] in shapes and shapes[-1] == [(512, 7)]


def test_mnist_builders_run():
    torch.manual_seed(0)
    assert B.mnist_prnmg(1, 1)(torch.randn(2, 1, 64, 64)).shape == (2, 1, 64, 64)
    assert B.mnist_unmg(10)(torch.randn(2, 1, 64, 64)).shape == (2, 10, 64, 64)


def test_oracle_reproduces_golden_vectors():
    """tests/golden/*.npz are this oracle's outputs on seeded inputs (the reference ships no fixtures):
    recomputing them pins the oracle against silent drift"""
    import importlib.util
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "make_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    assert len(mk.CASES) >= 4
    for name, fn in mk.CASES.items():
        stored = np.load(os.path.join(here, name))
        fresh = fn()
        assert set(stored.files) == set(fresh)
        for k in stored.files:
            a, b = stored[k], np.asarray(fresh[k])
            if a.dtype.kind in "iu":
                assert np.array_equal(a, b), (name, k)
            else:
                assert np.allclose(a, b, rtol=1e-6, atol=1e-6), (name, k)
