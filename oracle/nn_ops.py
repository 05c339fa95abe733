"""CPU ORACLE (test infrastructure, NOT product code) -- numpy restatement of the
upstream Torch7 `nn` operators that the reference's multigrid builders assemble.

PARITY UNPINNED: the reference (buttomnutstoast/Multigrid-Neural-Architectures)
ships no tests, golden vectors or runnable CPU path in this environment (no
Lua/Torch7; the arithmetic lives in un-vendored, unpinned luarocks packages
torch/nn (THNN), torch/cunn, cudnn.torch).  This file restates the published
THNN algorithms of those packages, anchored on the reference's own call sites:

  SpatialMaxPooling(2,2,2,2,0,0):ceil()   models/ilsvrc/rnmg.lua:57,201,214
  SpatialMaxPooling(3,3,2,2,1,1)          models/ilsvrc/rnmg.lua:183
  SpatialUpSamplingNearest(2)             models/ilsvrc/rnmg.lua:73
  JoinTable(2)                            models/ilsvrc/rnmg.lua:82
  SpatialConvolution(nIP,nOP,k,k,1,1,p,p) models/ilsvrc/rnmg.lua:26,36
  SpatialBatchNormalization(nOP[,eps])    models/ilsvrc/rnmg.lua:27,37; models/cifar/nmg.lua:23
  ReLU(true) / CAddTable(true)            models/ilsvrc/rnmg.lua:28,152
  Padding(1, nOP-nIP, 3)                  models/ilsvrc/rnmg.lua:16
  SpatialAveragePooling(r,r,r,r,0,0)      models/ilsvrc/rnmg.lua:175-177,282
  Linear / LogSoftMax / ClassNLLCriterion models/ilsvrc/rnmg.lua:283-285,325-329
  Sigmoid / BCECriterion                  models/mnist-cluttered/prnmg.mnist.lua:314,353-357
  optim.sgd                               models/basic_model.lua:64-66; pipelines/standard/train.lua:49-55

It is cross-checked op by op against PyTorch-CPU (direct descendants of the
same THNN sources) in tests/test_oracle.py.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg may import this module.

All tensors are NCHW numpy arrays (float64 by default for gradient checks).
Indices are 0-based (Lua's are 1-based: idx_lua = idx + 1).
"""
import numpy as np


# ----------------------------------------------------------------------------
# SpatialMaxPooling (THNN SpatialDilatedMaxPooling.c): window scan rows then
# cols, take v > max (or NaN), first max wins; index = y_in*W + x_in.
# ----------------------------------------------------------------------------
def _pool_out_size(H, k, s, p, ceil_mode):
    if ceil_mode:
        o = int(np.ceil(float(H + 2 * p - k) / s)) + 1
    else:
        o = int(np.floor(float(H + 2 * p - k) / s)) + 1
    if p > 0 or ceil_mode:
        # THNN: ensure that the last pooling starts inside the image
        if (o - 1) * s >= H + p:
            o -= 1
    return o


def maxpool_forward(x, k=2, s=2, p=0, ceil_mode=True):
    N, C, H, W = x.shape
    oH = _pool_out_size(H, k, s, p, ceil_mode)
    oW = _pool_out_size(W, k, s, p, ceil_mode)
    out = np.empty((N, C, oH, oW), dtype=x.dtype)
    idx = np.empty((N, C, oH, oW), dtype=np.int64)
    for oy in range(oH):
        y0 = oy * s - p
        y1 = min(y0 + k, H)
        y0 = max(y0, 0)
        for ox in range(oW):
            x0 = ox * s - p
            x1 = min(x0 + k, W)
            x0 = max(x0, 0)
            best = np.full((N, C), -np.inf, dtype=x.dtype)
            bidx = np.full((N, C), y0 * W + x0, dtype=np.int64)
            for yy in range(y0, y1):
                for xx in range(x0, x1):
                    v = x[:, :, yy, xx]
                    take = (v > best) | np.isnan(v)
                    best = np.where(take, v, best)
                    bidx = np.where(take, yy * W + xx, bidx)
            out[:, :, oy, ox] = best
            idx[:, :, oy, ox] = bidx
    return out, idx


def maxpool_backward(grad_out, idx, in_shape):
    N, C, H, W = in_shape
    gi = np.zeros((N, C, H * W), dtype=grad_out.dtype)
    flat_idx = idx.reshape(N, C, -1)
    flat_go = grad_out.reshape(N, C, -1)
    n_i, c_i = np.meshgrid(np.arange(N), np.arange(C), indexing="ij")
    for j in range(flat_idx.shape[2]):  # windows may overlap (3x3 s2): accumulate
        np.add.at(gi, (n_i, c_i, flat_idx[:, :, j]), flat_go[:, :, j])
    return gi.reshape(N, C, H, W)


# ----------------------------------------------------------------------------
# SpatialUpSamplingNearest(2) (THNN SpatialUpSamplingNearest.c)
# ----------------------------------------------------------------------------
def upsample_forward(x, r=2):
    return x.repeat(r, axis=2).repeat(r, axis=3)


def upsample_backward(grad_out, r=2):
    N, C, H, W = grad_out.shape
    return grad_out.reshape(N, C, H // r, r, W // r, r).sum(axis=(3, 5))


# ----------------------------------------------------------------------------
# SpatialAveragePooling(r,r,r,r,0,0) floor mode, divisor r*r
# ----------------------------------------------------------------------------
def avgpool_forward(x, k, s=None):
    s = s or k
    N, C, H, W = x.shape
    oH = (H - k) // s + 1
    oW = (W - k) // s + 1
    out = np.zeros((N, C, oH, oW), dtype=x.dtype)
    for dy in range(k):
        for dx in range(k):
            out += x[:, :, dy:dy + (oH - 1) * s + 1:s, dx:dx + (oW - 1) * s + 1:s]
    return out / (k * k)


def avgpool_backward(grad_out, in_shape, k, s=None):
    s = s or k
    N, C, H, W = in_shape
    oH, oW = grad_out.shape[2:]
    gi = np.zeros(in_shape, dtype=grad_out.dtype)
    for dy in range(k):
        for dx in range(k):
            gi[:, :, dy:dy + (oH - 1) * s + 1:s, dx:dx + (oW - 1) * s + 1:s] += grad_out
    return gi / (k * k)


# ----------------------------------------------------------------------------
# SpatialConvolution (THNN SpatialConvolutionMM.c = im2col + GEMM):
# cross-correlation, y = W * x + b, weight [Cout, Cin, kH, kW]
# ----------------------------------------------------------------------------
def _im2col(x, k, s, p):
    N, C, H, W = x.shape
    oH = (H + 2 * p - k) // s + 1
    oW = (W + 2 * p - k) // s + 1
    xp = np.zeros((N, C, H + 2 * p, W + 2 * p), dtype=x.dtype)
    xp[:, :, p:p + H, p:p + W] = x
    cols = np.empty((N, C, k, k, oH, oW), dtype=x.dtype)
    for dy in range(k):
        for dx in range(k):
            cols[:, :, dy, dx] = xp[:, :, dy:dy + (oH - 1) * s + 1:s, dx:dx + (oW - 1) * s + 1:s]
    return cols, oH, oW


def conv_forward(x, w, b=None, s=1, p=None):
    Cout, Cin, k, _ = w.shape
    if p is None:
        p = 0 if k == 1 else 1
    cols, oH, oW = _im2col(x, k, s, p)
    N = x.shape[0]
    y = np.einsum("ok,nkp->nop", w.reshape(Cout, -1), cols.reshape(N, Cin * k * k, oH * oW))
    y = y.reshape(N, Cout, oH, oW)
    if b is not None:
        y = y + b.reshape(1, -1, 1, 1)
    return y


def conv_backward(x, w, grad_out, s=1, p=None):
    """returns (grad_input, grad_weight, grad_bias) -- accGradParameters with scale 1"""
    Cout, Cin, k, _ = w.shape
    if p is None:
        p = 0 if k == 1 else 1
    N, _, H, W = x.shape
    cols, oH, oW = _im2col(x, k, s, p)
    go = grad_out.reshape(N, Cout, oH * oW)
    gw = np.einsum("nop,nkp->ok", go, cols.reshape(N, Cin * k * k, oH * oW)).reshape(w.shape)
    gb = go.sum(axis=(0, 2))
    gcols = np.einsum("ok,nop->nkp", w.reshape(Cout, -1), go).reshape(N, Cin, k, k, oH, oW)
    gxp = np.zeros((N, Cin, H + 2 * p, W + 2 * p), dtype=x.dtype)
    for dy in range(k):
        for dx in range(k):
            gxp[:, :, dy:dy + (oH - 1) * s + 1:s, dx:dx + (oW - 1) * s + 1:s] += gcols[:, :, dy, dx]
    return gxp[:, :, p:p + H, p:p + W], gw, gb


# ----------------------------------------------------------------------------
# SpatialFullConvolution(nIP,nOP,2,2,2,2,0,0) (unmg.lua:36): weight [Cin,Cout,2,2]
# ----------------------------------------------------------------------------
def upconv2x2_forward(x, w, b=None):
    N, Cin, H, W = x.shape
    Cout = w.shape[1]
    y = np.zeros((N, Cout, 2 * H, 2 * W), dtype=x.dtype)
    for dy in range(2):
        for dx in range(2):
            y[:, :, dy::2, dx::2] = np.einsum("nchw,co->nohw", x, w[:, :, dy, dx])
    if b is not None:
        y = y + b.reshape(1, -1, 1, 1)
    return y


def upconv2x2_backward(x, w, grad_out):
    gx = np.zeros_like(x)
    gw = np.zeros_like(w)
    for dy in range(2):
        for dx in range(2):
            g = grad_out[:, :, dy::2, dx::2]
            gx += np.einsum("nohw,co->nchw", g, w[:, :, dy, dx])
            gw[:, :, dy, dx] = np.einsum("nchw,nohw->co", x, g)
    return gx, gw, grad_out.sum(axis=(0, 2, 3))


# ----------------------------------------------------------------------------
# SpatialBatchNormalization (THNN BatchNormalization.c)
# ----------------------------------------------------------------------------
def bn_forward_train(x, gamma, beta, eps=1e-5, running_mean=None, running_var=None, momentum=0.1):
    """returns y, save_mean, save_invstd; updates running stats in place."""
    n = x.shape[0] * x.shape[2] * x.shape[3]
    xd = x.astype(np.float64)  # THNN CPU accumulates in double
    mean = xd.mean(axis=(0, 2, 3))
    var = ((xd - mean.reshape(1, -1, 1, 1)) ** 2).sum(axis=(0, 2, 3)) / n  # biased
    invstd = 1.0 / np.sqrt(var + eps)
    if running_mean is not None:
        running_mean *= (1 - momentum)
        running_mean += momentum * mean
        unbiased = var * n / max(n - 1, 1)
        running_var *= (1 - momentum)
        running_var += momentum * unbiased
    xhat = (xd - mean.reshape(1, -1, 1, 1)) * invstd.reshape(1, -1, 1, 1)
    y = xhat * gamma.reshape(1, -1, 1, 1) + beta.reshape(1, -1, 1, 1)
    return y.astype(x.dtype), mean, invstd


def bn_forward_eval(x, gamma, beta, running_mean, running_var, eps=1e-5):
    invstd = 1.0 / np.sqrt(running_var + eps)
    y = (x - running_mean.reshape(1, -1, 1, 1)) * (invstd * gamma).reshape(1, -1, 1, 1) \
        + beta.reshape(1, -1, 1, 1)
    return y.astype(x.dtype)


def bn_backward_train(x, grad_out, gamma, mean, invstd):
    """returns (grad_input, grad_gamma, grad_beta)"""
    n = x.shape[0] * x.shape[2] * x.shape[3]
    xhat = (x - mean.reshape(1, -1, 1, 1)) * invstd.reshape(1, -1, 1, 1)
    dgamma = (grad_out * xhat).sum(axis=(0, 2, 3))
    dbeta = grad_out.sum(axis=(0, 2, 3))
    gi = (grad_out - dbeta.reshape(1, -1, 1, 1) / n - xhat * dgamma.reshape(1, -1, 1, 1) / n) \
        * (gamma * invstd).reshape(1, -1, 1, 1)
    return gi.astype(x.dtype), dgamma, dbeta


# ----------------------------------------------------------------------------
# pointwise / table ops
# ----------------------------------------------------------------------------
def relu_forward(x):
    return np.maximum(x, 0)


def relu_backward(y, grad_out):
    return grad_out * (y > 0)


def pad_channels(x, nOP):
    """nn.Padding(1, nOP-nIP, 3) on a 4-D batch: zeros appended AFTER the last channel."""
    N, C, H, W = x.shape
    if nOP == C:
        return x
    out = np.zeros((N, nOP, H, W), dtype=x.dtype)
    out[:, :C] = x
    return out


def join_channels(xs):
    return np.concatenate(xs, axis=1)


# ----------------------------------------------------------------------------
# ResampleConcat (models/ilsvrc/rnmg.lua:41-89; isDrop variant prnmg.mnist.lua:44-92)
# channel order: finer (max-pooled) | same | coarser (nearest-upsampled)
# ----------------------------------------------------------------------------
def resample_concat_forward(xs, is_drop=False):
    n = len(xs) - 1 if is_drop else len(xs)
    outs, saved = [], []
    for i in range(n):
        parts, idx = [], None
        if i - 1 >= 0:
            pooled, idx = maxpool_forward(xs[i - 1])
            parts.append(pooled)
        parts.append(xs[i])
        if i + 1 < n:
            parts.append(upsample_forward(xs[i + 1]))
        outs.append(join_channels(parts))
        saved.append(idx)
    return outs, saved


def resample_concat_backward(xs, saved_idx, grad_outs, is_drop=False):
    """ConcatTable backward = SUM of branch gradInputs per table entry."""
    n = len(xs) - 1 if is_drop else len(xs)
    gxs = [np.zeros_like(x) for x in xs]
    for i in range(n):
        g = grad_outs[i]
        c0 = 0
        if i - 1 >= 0:
            c = xs[i - 1].shape[1]
            gxs[i - 1] += maxpool_backward(g[:, c0:c0 + c], saved_idx[i], xs[i - 1].shape)
            c0 += c
        c = xs[i].shape[1]
        gxs[i] += g[:, c0:c0 + c]
        c0 += c
        if i + 1 < n:
            c = xs[i + 1].shape[1]
            gxs[i + 1] += upsample_backward(g[:, c0:c0 + c])
    return gxs


# ----------------------------------------------------------------------------
# one plain mg-conv stage and one residual mg unit, forward + backward
# (models/cifar/nmg.lua:31-86; models/ilsvrc/rnmg.lua:91-159)
# params per scale: dict(w, b, gamma, beta, rm, rv)
# ----------------------------------------------------------------------------
def mg_stage_forward(xs, params, eps=1e-5, relu=True, is_drop=False, train=True):
    cats, idxs = resample_concat_forward(xs, is_drop)
    ys, cache = [], []
    for i, cat in enumerate(cats):
        p = params[i]
        conv = conv_forward(cat, p["w"], p["b"])
        if train:
            bn, mean, invstd = bn_forward_train(conv, p["gamma"], p["beta"], eps, p.get("rm"), p.get("rv"))
        else:
            bn, mean, invstd = bn_forward_eval(conv, p["gamma"], p["beta"], p["rm"], p["rv"], eps), None, None
        y = relu_forward(bn) if relu else bn
        ys.append(y)
        cache.append((cat, conv, mean, invstd, y))
    return ys, (xs, idxs, cache, relu, is_drop)


def mg_stage_backward(grad_ys, params, ctx):
    xs, idxs, cache, relu, is_drop = ctx
    gcats, gparams = [], []
    for i, gy in enumerate(grad_ys):
        cat, conv, mean, invstd, y = cache[i]
        p = params[i]
        g = relu_backward(y, gy) if relu else gy
        gconv, dgamma, dbeta = bn_backward_train(conv, g, p["gamma"], mean, invstd)
        gcat, gw, gb = conv_backward(cat, p["w"], gconv)
        gcats.append(gcat)
        gparams.append(dict(w=gw, b=gb, gamma=dgamma, beta=dbeta))
    return resample_concat_backward(xs, idxs, gcats, is_drop), gparams


def mg_resunit_forward(xs, params1, params2, eps=1e-5, final_relu=True, is_drop=False):
    """ReLU(BN(mg(ReLU(BN(mg(x))))) + Shortcut(x)); zero-padded identity shortcut."""
    h, ctx1 = mg_stage_forward(xs, params1, eps, relu=True, is_drop=is_drop)
    z, ctx2 = mg_stage_forward(h, params2, eps, relu=False)
    outs = []
    for i in range(len(z)):
        s = z[i] + pad_channels(xs[i], z[i].shape[1])
        outs.append(relu_forward(s) if final_relu else s)
    return outs, (ctx1, ctx2, outs, final_relu)


def mg_resunit_backward(grad_outs, params1, params2, ctx):
    ctx1, ctx2, outs, final_relu = ctx
    xs = ctx1[0]
    gz = [relu_backward(outs[i], grad_outs[i]) if final_relu else grad_outs[i] for i in range(len(outs))]
    gh, gp2 = mg_stage_backward(gz, params2, ctx2)
    gx, gp1 = mg_stage_backward(gh, params1, ctx1)
    for i in range(len(gz)):
        gx[i] = gx[i] + gz[i][:, :xs[i].shape[1]]
    return gx, gp1, gp2


# ----------------------------------------------------------------------------
# head + criteria
# ----------------------------------------------------------------------------
def linear_forward(x, w, b):
    return x @ w.T + b


def linear_backward(x, w, grad_out):
    return grad_out @ w, grad_out.T @ x, grad_out.sum(axis=0)


def logsoftmax_forward(x):
    m = x.max(axis=1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(axis=1, keepdims=True))


def logsoftmax_backward(y, grad_out):
    return grad_out - np.exp(y) * grad_out.sum(axis=1, keepdims=True)


def nll_forward(logp, target):
    """ClassNLLCriterion, sizeAverage=true, 0-based targets here."""
    return -logp[np.arange(len(target)), target].mean()


def nll_backward(logp, target):
    g = np.zeros_like(logp)
    g[np.arange(len(target)), target] = -1.0 / len(target)
    return g


def sigmoid_forward(x):
    return 1.0 / (1.0 + np.exp(-x))


def bce_forward(p, t, eps=1e-12):
    """BCECriterion, sizeAverage over all elements, eps 1e-12."""
    return -(np.log(p + eps) * t + np.log(1 - p + eps) * (1 - t)).mean()


def bce_backward(p, t, eps=1e-12):
    return -(t - p) / ((1 - p + eps) * (p + eps)) / p.size


# ----------------------------------------------------------------------------
# optim.sgd (dampening 0, nesterov off): g += wd*w; first call v=g else v=mu*v+g; w -= lr*v
# ----------------------------------------------------------------------------
def sgd_step(w, g, state, lr, momentum=0.9, wd=0.0):
    g = g + wd * w
    if momentum != 0:
        if "v" not in state:
            state["v"] = g.copy()
        else:
            state["v"] = momentum * state["v"] + g
        g = state["v"]
    return w - lr * g
