"""CPU ORACLE (test infrastructure, NOT product code) -- a Torch7-`nn`-shaped
container/module vocabulary on top of PyTorch-CPU, so that oracle/builders.py can
restate the reference's Lua model builders line by line.

PARITY UNPINNED (see oracle/nn_ops.py header): Torch7 itself cannot run here.
PyTorch's CPU kernels descend from the same THNN sources; pooling tie-break/index
and nearest-upsample conventions are checked against the numpy restatement in
tests/test_oracle.py.

Torch7 semantics restated here (upstream torch/nn, un-vendored):
  nn.Sequential, nn.ConcatTable (same input to every branch, table out; backward
  sums), nn.ParallelTable (i-th module on i-th entry), nn.SelectTable(i) (1-based),
  nn.JoinTable(2) (channel concat of a table of 4-D tensors), nn.FlattenTable,
  nn.CAddTable, nn.MapTable, nn.Identity, nn.Padding(1, pad, 3) (zero channels
  appended after the last channel), nn.View(-1, n).
Tables are Python lists.  autograd supplies every backward.
"""
import math
import torch
import torch.nn as tnn
import torch.nn.functional as F


class Sequential(tnn.Sequential):
    def add(self, m):
        self.append(m)
        return self


class ConcatTable(tnn.Module):
    def __init__(self):
        super().__init__()
        self.mods = tnn.ModuleList()

    def add(self, m):
        self.mods.append(m)
        return self

    def forward(self, x):
        return [m(x) for m in self.mods]


class ParallelTable(ConcatTable):
    def forward(self, x):
        return [m(x[i]) for i, m in enumerate(self.mods)]


class MapTable(tnn.Module):
    def __init__(self, m):
        super().__init__()
        self.m = m

    def forward(self, x):
        return [self.m(e) for e in x]


class SelectTable(tnn.Module):
    def __init__(self, i):  # 1-based like Lua
        super().__init__()
        self.i = i

    def forward(self, x):
        return x[self.i - 1]


class JoinTable(tnn.Module):
    def __init__(self, dim):  # dim=2 on 4-D batches = channels
        super().__init__()
        assert dim == 2

    def forward(self, x):
        return torch.cat(list(x), dim=1)


class FlattenTable(tnn.Module):
    def forward(self, x):
        out = []

        def rec(t):
            if isinstance(t, (list, tuple)):
                for e in t:
                    rec(e)
            else:
                out.append(t)
        rec(x)
        return out


class CAddTable(tnn.Module):
    def __init__(self, inplace=False):
        super().__init__()

    def forward(self, x):
        s = x[0]
        for e in x[1:]:
            s = s + e
        return s


class Identity(tnn.Identity):
    pass


class Padding(tnn.Module):
    """nn.Padding(1, pad, 3): per-sample dim 1 (= channels), pad>0 appends zeros."""

    def __init__(self, dim, pad, nInputDim):
        super().__init__()
        assert dim == 1 and nInputDim == 3 and pad > 0
        self.pad = pad

    def forward(self, x):
        return F.pad(x, (0, 0, 0, 0, 0, self.pad))


class View(tnn.Module):
    def __init__(self, *shape):
        super().__init__()
        self.shape = shape

    def forward(self, x):
        return x.reshape(*self.shape)


class ConcatUnet(tnn.Module):
    """layers/ConcatUnet.lua:1-37: {shortcut{t1..tn}, subnet{p1..pm}} -> {{t1,p1},...,{tn[,pn]}}"""

    def forward(self, x):
        assert len(x) == 2 and len(x[0]) >= len(x[1])
        out = []
        for i in range(len(x[0])):
            e = [x[0][i]]
            if i < len(x[1]):
                e.append(x[1][i])
            out.append(e)
        return out


# ---- leaf modules with Torch7 constructors / default reset() --------------------
def SpatialConvolution(nIP, nOP, kW, kH, dW=1, dH=1, padW=0, padH=0):
    m = tnn.Conv2d(nIP, nOP, (kH, kW), (dH, dW), (padH, padW), bias=True)
    # torch7 default reset(): uniform(-1/sqrt(kW*kH*nIP), +) for weight and bias
    stdv = 1.0 / math.sqrt(kW * kH * nIP)
    with torch.no_grad():
        m.weight.uniform_(-stdv, stdv)
        m.bias.uniform_(-stdv, stdv)
    return m


def SpatialFullConvolution(nIP, nOP, kW, kH, dW, dH, padW=0, padH=0):
    m = tnn.ConvTranspose2d(nIP, nOP, (kH, kW), (dH, dW), (padH, padW), bias=True)
    stdv = 1.0 / math.sqrt(kW * kH * nIP)
    with torch.no_grad():
        m.weight.uniform_(-stdv, stdv)
        m.bias.uniform_(-stdv, stdv)
    return m


def SpatialBatchNormalization(nOP, eps=1e-5, momentum=0.1):
    m = tnn.BatchNorm2d(nOP, eps=eps, momentum=momentum, affine=True)
    with torch.no_grad():  # torch7 reset(): weight ~ U(0,1), bias 0, rm 0, rv 1
        m.weight.uniform_(0, 1)
        m.bias.zero_()
    return m


def ReLU(inplace=False):
    return tnn.ReLU(inplace=False)


class _MaxPool(tnn.MaxPool2d):
    def ceil(self):
        self.ceil_mode = True
        return self


def SpatialMaxPooling(kW, kH, dW, dH, padW=0, padH=0):
    return _MaxPool((kH, kW), (dH, dW), (padH, padW))


def SpatialAveragePooling(kW, kH, dW, dH, padW=0, padH=0):
    return tnn.AvgPool2d((kH, kW), (dH, dW), (padH, padW))


def SpatialUpSamplingNearest(r):
    return tnn.Upsample(scale_factor=r, mode="nearest")


def Linear(nIn, nOut):
    m = tnn.Linear(nIn, nOut)
    stdv = 1.0 / math.sqrt(nIn)
    with torch.no_grad():
        m.weight.uniform_(-stdv, stdv)
        m.bias.uniform_(-stdv, stdv)
    return m


LogSoftMax = lambda: tnn.LogSoftmax(dim=1)
Sigmoid = tnn.Sigmoid


def find_modules(model, cls):
    return [m for m in model.modules() if isinstance(m, cls)]


def conv_init_msr_fanout(model):
    """ConvInit in models/ilsvrc/rnmg.lua:288-294: N(0, sqrt(2/(kW*kH*nOutputPlane))), bias 0"""
    for v in find_modules(model, tnn.Conv2d):
        n = v.kernel_size[0] * v.kernel_size[1] * v.out_channels
        with torch.no_grad():
            v.weight.normal_(0, math.sqrt(2.0 / n))
            v.bias.zero_()


def bn_init(model):
    """BNInit in models/ilsvrc/rnmg.lua:295-300: gamma=1, beta=0"""
    for v in find_modules(model, tnn.BatchNorm2d):
        with torch.no_grad():
            v.weight.fill_(1)
            v.bias.zero_()


def linear_bias_zero(model):
    for v in find_modules(model, tnn.Linear):
        with torch.no_grad():
            v.bias.zero_()
