"""Generates the golden vectors under tests/golden/ from the numpy oracle (oracle/nn_ops.py, fp64).

The reference ships no tests, fixtures or runnable CPU path, so these vectors are the
oracle's own outputs on seeded inputs, committed so that (a) the oracle cannot drift unnoticed
(tests/test_oracle.py::test_oracle_reproduces_golden_vectors recomputes them on CPU) and (b) the CUDA path is
held to fixed numbers (tests/test_kernels_gpu.py::test_golden_*).  Inputs and weights are bf16-representable
so that the bf16 kernels see exactly these values.

    python tests/golden/make_golden.py        # rewrites the .npz files (deterministic)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import nn_ops as O  # noqa: E402


def bf16(a):
    return torch.from_numpy(np.asarray(a, dtype=np.float32)).to(torch.bfloat16).to(torch.float64).numpy()


def mgconv_case(seed, N, cs, H, Cout, k):
    """one multigrid convolution: y = conv_k(cat[maxpool2x2_ceil(finer), same, upsample2(coarser)]) and its backward
    (models/ilsvrc/rnmg.lua:41-89 + 26)"""
    r = np.random.default_rng(seed)
    f, s, c = bf16(r.standard_normal((N, cs[0], 2 * H, 2 * H))), bf16(r.standard_normal((N, cs[1], H, H))), bf16(r.standard_normal((N, cs[2], H // 2, H // 2)))
    w, b = bf16(r.standard_normal((Cout, sum(cs), k, k)) * 0.2), bf16(r.standard_normal(Cout) * 0.1)
    g = bf16(r.standard_normal((N, Cout, H, H)))
    pad = 0 if k == 1 else 1
    pooled, idx = O.maxpool_forward(f)
    cat = np.concatenate([pooled, s, O.upsample_forward(c)], axis=1)
    y = O.conv_forward(cat, w, b, 1, pad)
    gcat, gw, gb = O.conv_backward(cat, w, g, 1, pad)
    # gradients w.r.t. the three grids: through the pool arg-max, identity, 2x2 block sum
    gf = O.maxpool_backward(gcat[:, :cs[0]], idx, f.shape)
    gs = gcat[:, cs[0]:cs[0] + cs[1]]
    gc = O.upsample_backward(gcat[:, cs[0] + cs[1]:])
    return dict(finer=f, same=s, coarser=c, weight=w, bias=b, grad_out=g, k=np.int64(k), pool_argmax=idx.astype(np.int64), y=y,
                grad_cat=gcat, grad_weight=gw, grad_bias=gb, grad_finer=gf, grad_same=gs, grad_coarser=gc)


def bn_case(seed, N, C, H, Cs, eps):
    """SpatialBatchNormalization (train) -> CAddTable with a zero-padded shortcut -> ReLU, and backward
    (models/ilsvrc/rnmg.lua:13-20, 27, 140-154)"""
    r = np.random.default_rng(seed)
    x = bf16(r.standard_normal((N, C, H, H)) * 2 + 1)
    gamma, beta = r.random(C) + 0.5, r.standard_normal(C) * 0.1
    rm, rv = np.zeros(C), np.ones(C)
    y, mean, invstd = O.bn_forward_train(x, gamma, beta, eps, rm, rv)
    sc = bf16(r.standard_normal((N, Cs, H, H)))
    out = O.relu_forward(y + O.pad_channels(sc, C))
    go = bf16(r.standard_normal((N, C, H, H)))
    d = O.relu_backward(out, go)
    gx, dgamma, dbeta = O.bn_backward_train(x, d, gamma, mean, invstd)
    return dict(x=x, gamma=gamma, beta=beta, eps=np.float64(eps), shortcut=sc, out=out, pooled=O.maxpool_forward(out)[0], running_mean=rm, running_var=rv,
                save_mean=mean, save_invstd=invstd, grad_out=go, grad_x=gx, grad_gamma=dgamma, grad_beta=dbeta, grad_shortcut=d[:, :Cs])


CASES = {
    "mgconv_3scale_k3.npz": lambda: mgconv_case(11, 2, (16, 8, 8), 14, 24, 3),
    "mgconv_3scale_k1.npz": lambda: mgconv_case(12, 2, (8, 12, 4), 4, 10, 1),
    "bn_shortcut_relu.npz": lambda: bn_case(21, 3, 12, 6, 8, 1e-5),
    "bn_shortcut_relu_odd.npz": lambda: bn_case(22, 2, 20, 7, 20, 1e-3),
}

if __name__ == "__main__":
    for name, fn in CASES.items():
        d = fn()
        np.savez_compressed(os.path.join(HERE, name), **{k: (v.astype(np.float32) if v.dtype == np.float64 and k not in ("eps",) else v) for k, v in d.items()})
        print(name, sum(v.nbytes for v in d.values()) // 1024, "KiB raw")
