"""helpers shared by the GPU parity tests: grids from numpy, oracle <-> product weight transfer"""
import ctypes as C
import numpy as np
import torch

from mgconv import ffi
from mgconv.ffi import mg_grid, mg_conv_desc, mg_grad_src, ptr

TDT = {ffi.MG_F32: torch.float32, ffi.MG_BF16: torch.bfloat16}
# tolerances of BASELINE.json north_star: "rel 2e-2" (bf16, fp32 accumulate), "fp32 mode rel 1e-4"
TOL = {ffi.MG_F32: 1e-4, ffi.MG_BF16: 2e-2}


def cpad(c):
    return (c + 7) // 8 * 8


def bf16_round(a):
    """numpy fp64/fp32 array rounded to bf16-representable values (so both sides see the same inputs)"""
    return torch.from_numpy(np.asarray(a, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy().astype(np.float64)


class Grid:
    """NHWC device tensor + its mg_grid view"""

    def __init__(self, dtype, N, C, H, W, nchw=None, Cp=None):
        self.N, self.C, self.H, self.W = N, C, H, W
        self.Cp = Cp or cpad(C)
        self.t = torch.zeros((N, H, W, self.Cp), dtype=TDT[dtype], device="cuda")
        if nchw is not None:
            self.t[..., :C] = torch.from_numpy(np.ascontiguousarray(nchw)).to("cuda").permute(0, 2, 3, 1).to(TDT[dtype])
        self.scale = self.shift = None
        self.relu = 0

    def affine(self, scale, shift, relu):
        self.scale = torch.zeros(self.Cp, dtype=torch.float32, device="cuda")
        self.shift = torch.zeros(self.Cp, dtype=torch.float32, device="cuda")
        self.scale[:self.C] = torch.from_numpy(np.asarray(scale, dtype=np.float32)).cuda()
        self.shift[:self.C] = torch.from_numpy(np.asarray(shift, dtype=np.float32)).cuda()
        self.relu = int(relu)
        return self

    def g(self):
        return mg_grid(self.t.data_ptr(), None if self.scale is None else self.scale.data_ptr(),
                       None if self.shift is None else self.shift.data_ptr(), self.relu,
                       self.N, self.H, self.W, self.C, self.Cp)

    def nchw(self):
        torch.cuda.synchronize()
        return self.t[..., :self.C].permute(0, 3, 1, 2).float().cpu().numpy().astype(np.float64)

    def pad_channels(self):
        torch.cuda.synchronize()
        return self.t[..., self.C:].float().cpu().numpy()


def conv_desc(segs, modes, k, stride, pad, Cout, H, W):
    d = mg_conv_desc()
    d.n_seg = len(segs)
    for i, (s, m) in enumerate(zip(segs, modes)):
        d.seg[i] = s.g()
        d.seg_mode[i] = m
    d.ksize, d.stride, d.pad, d.Cout, d.H, d.W = k, stride, pad, Cout, H, W
    return d


def dev(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype).cuda()


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def max_rel(a, b):
    """max |a-b| relative to the largest magnitude of the reference tensor"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def elem_rel(a, b, floor_frac=0.02):
    """per-ELEMENT relative error max |a - b| / max(|b|, floor), floor = floor_frac x the largest magnitude of the reference
    (elements smaller than the floor are held to an absolute error of tol x floor): the north-star's `rel` read element-wise"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    floor = max(floor_frac * np.abs(b).max(), 1e-30)
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor)).max())


def copy_params_from_oracle(omodel, model, bf16_weights=False):
    """oracle (torch.nn) parameters/buffers -> product modules, matched in construction order.
    bf16_weights: first round the oracle's conv / linear weights to bf16-representable values, so
    that the bf16 tensor-core path (which packs its operands to bf16) sees identical weights."""
    import torch.nn as tnn
    olist = [m for m in omodel.modules() if isinstance(m, (tnn.Conv2d, tnn.BatchNorm2d, tnn.Linear))]
    if bf16_weights:
        with torch.no_grad():
            for o in olist:
                if not isinstance(o, tnn.BatchNorm2d):
                    o.weight.copy_(o.weight.to(torch.bfloat16).to(o.weight.dtype))
    plist = [m for m in model.listModules() if m.own_parameters()]
    assert len(olist) == len(plist), (len(olist), len(plist))
    for o, p in zip(olist, plist):
        assert tuple(o.weight.shape) == tuple(p.weight.shape), (type(o).__name__, p.typename, o.weight.shape, p.weight.shape)
        p.weight.copy_(o.weight.detach())
        p.bias.copy_(o.bias.detach())
        if isinstance(o, tnn.BatchNorm2d):
            p.running_mean.copy_(o.running_mean)
            p.running_var.copy_(o.running_var)
    return olist, plist


def emulate_bf16_storage(omodel):
    """bf16 mode of the product = bf16 *storage* of every activation / gradient tensor, fp32
    accumulation and fp32 BatchNorm arithmetic.  The fp64 oracle is given the same storage points
    (forward hooks that round to bf16; autograd rounds the gradient at the same places), so that
    ReLU masks and pool arg-maxima are decided on identical values and the remaining difference is
    accumulation order only."""
    import torch.nn as tnn
    from oracle import t7nn

    def rnd(_m, _i, o):
        return o.to(torch.bfloat16).to(o.dtype)
    for m in omodel.modules():
        if isinstance(m, (tnn.Conv2d, tnn.Linear, tnn.ReLU, tnn.AvgPool2d, t7nn.CAddTable)):
            m.register_forward_hook(rnd)
    return omodel


def new_sums(n):
    """buffer of n deterministic sums (mg_sum: two int64 limbs each, include/mgconv.h), zeroed"""
    import torch
    return torch.zeros(2 * n, dtype=torch.int64, device="cuda")


def sums_value(t):
    """mg_sum buffer -> float64 tensor of the values (hi * 2^-10 + lo * 2^-54)"""
    import torch
    v = t.view(-1, 2).to(torch.float64)
    return v[:, 0] * 2.0 ** -10 + v[:, 1] * 2.0 ** -54
