"""optim.sgd of the reference's step driver (models/basic_model.lua:64-66; state built in
pipelines/standard/train.lua:49-55) as one fused pass over the flat parameter vector:
    g += wd*w ; v = (first call ? g : mu*v + g) ; w -= lr*v        (dampening 0, nesterov off)
"""
import torch

from .engine import criterion_ctx
from .ffi import ptr


def sgd(feval, x, state):
    lr = state.get("learningRate", 1e-3)
    lrd = state.get("learningRateDecay", 0.0)
    wd = state.get("weightDecay", 0.0)
    mom = state.get("momentum", 0.0)
    damp = state.get("dampening", mom if "dampening" not in state and False else 0.0)
    if damp != 0.0:
        raise NotImplementedError("optim.sgd dampening != 0 is never used by the reference (train.lua:53)")
    state["evalCounter"] = state.get("evalCounter", 0)
    fx, dfdx = feval(x)
    clr = lr / (1 + state["evalCounter"] * lrd)
    first = "dfdx" not in state
    if first:
        state["dfdx"] = torch.zeros_like(x)
    ctx = criterion_ctx(x)
    ctx.call("mg_sgd_step", ptr(x), ptr(dfdx), ptr(state["dfdx"]), x.numel(), float(clr), float(mom), float(wd), int(first))
    state["evalCounter"] += 1
    return x, [fx]
