"""utils/modelfuncs.lua: parameter initialisers keyed on `findModules(<backend>.<type>)` and the builder's smoke test.

The module vocabulary of mgconv.nn registers convolutions under the type name the reference's builders search for
('cudnn.SpatialConvolution', ilsvrc/rnmg.lua:302); `backend` is accepted for either spelling, as cudnn.convert() would
leave a model (init code of the reference anticipates both, ilsvrc/rnmg.lua:302-305)."""
import math

import torch


def _convs(model, backend):
    assert backend in ("nn", "cudnn")                                            # modelfuncs.lua:4
    found = model.findModules(backend + ".SpatialConvolution")
    other = "cudnn" if backend == "nn" else "nn"
    return found or model.findModules(other + ".SpatialConvolution")


def MSRinit(model, backend="cudnn", pow=0.5):
    """modelfuncs.lua:3-11: weight ~ N(0, (2 / (kW*kH*nInputPlane))^pow) -- fan-IN, with the exponent as a parameter
    (pow = 0.5 is He et al.); bias 0.  (The ImageNet / CIFAR builders use their own fan-OUT ConvInit: builders.MSRinit.)"""
    for v in _convs(model, backend):
        n = v.kW * v.kH * v.nInputPlane
        v.weight.normal_(0, math.pow(2.0 / n, pow))
        if getattr(v, "bias", None) is not None:
            v.bias.zero_()


def XAVinit(model, backend="cudnn", const=1.0, sqrt=6.0):
    """modelfuncs.lua:13-22: weight ~ U(-val, val), val = const * sqrt(sqrt / (nInputPlane + nOutputPlane)); bias 0"""
    for v in _convs(model, backend):
        val = const * math.sqrt(sqrt / (v.nInputPlane + v.nOutputPlane))
        v.weight.uniform_(-val, val)
        if getattr(v, "bias", None) is not None:
            v.bias.zero_()


def GAUSSinit(model, backend="cudnn", mean=0.0, stddev=0.01):
    """modelfuncs.lua:24-30: weight ~ N(mean, stddev); bias 0"""
    for v in _convs(model, backend):
        v.weight.normal_(mean, stddev)
        if getattr(v, "bias", None) is not None:
            v.bias.zero_()


def FCinit(model):
    """modelfuncs.lua:33-37"""
    for v in model.findModules("nn.Linear"):
        v.bias.zero_()


def BNinit(model, backend="nn", w_const=1.0, bias_const=0.0):
    """modelfuncs.lua:39-46"""
    found = model.findModules(backend + ".SpatialBatchNormalization") or model.findModules("nn.SpatialBatchNormalization")
    for v in found:
        v.weight.fill_(w_const)
        v.bias.fill_(bias_const)


def DisableBias(model, backend="cudnn"):
    """modelfuncs.lua:48-54: the reference drops bias / gradBias of every convolution (they are absorbed by the BatchNorm that
    follows).  Here the storage stays (the flat parameter vector keeps its layout) but the bias is pinned to zero."""
    for v in _convs(model, backend):
        v.bias.zero_()
        v.noBias = True


def testModel(model, imageSize=32, device=None):
    """modelfuncs.lua:56-63: one forward / backward on a random image, shapes printed, parameters re-drawn (`model:reset()`).
    The reference runs it on the CPU (`model:float()`); the multigrid hot path has no CPU implementation, so the model is
    moved to the GPU instead."""
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    model.cuda(dev.index)
    x = torch.randn(1, 3, imageSize, imageSize, device=dev)
    out = model.forward(x)
    print("forward output", [tuple(o.shape) for o in (out if isinstance(out, (list, tuple)) else [out])])
    gi = model.backward(x, out)
    print("backward output", None if gi is None else [tuple(g.shape) for g in (gi if isinstance(gi, (list, tuple)) else [gi])])
    for m in model.listModules():
        if hasattr(m, "reset"):
            m.reset()
    return out
