"""micro-benchmark of single multigrid convolutions (R-MG-34 shapes, B=256) through the C ABI:
forward / dgrad / wgrad device time with CUDA events; the tuning harness for the tcgen05 kernels."""
import ctypes as C
import os
import sys
import json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multigrid-neural-architectures_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from mgconv import ffi
from mgconv.ffi import ptr, MG_SEG_SAME, MG_SEG_UP
from util import Grid, conv_desc

SHAPES = {
    # name: (H, [(C, mode)], Cout, k)        R-MG-34 at B = 256
    "b1g1": (56, [(64, "s"), (32, "u")], 64, 3),
    "b1g2": (28, [(64, "s"), (32, "s"), (16, "u")], 32, 3),
    "b2g1": (28, [(128, "s"), (64, "u")], 128, 3),
    "b2g2": (14, [(128, "s"), (64, "s"), (32, "u")], 64, 3),
    "b3g1": (14, [(256, "s"), (128, "u")], 256, 3),
    "b3g2": (7, [(256, "s"), (128, "s")], 128, 3),
    "b4": (7, [(512, "s")], 512, 3),
    # ImageNet stem: cudnn.SpatialConvolution(3, 64, 7,7, 2,2, 3,3) on the 224x224 image (ilsvrc/rnmg.lua:180); (H, segs, Cout, k, stride, pad)
    "stem": (224, [(3, "s")], 64, 7, 2, 3),
}


def run(name, N=256, iters=20, which=("fwd", "dgrad", "wgrad")):
    H, segs, Cout, k = SHAPES[name][:4]
    stride = SHAPES[name][4] if len(SHAPES[name]) > 4 else 1
    ctx = ffi.Context(0, torch.cuda.current_stream().cuda_stream, ffi.MG_BF16)
    ctx.set_tuning(ffi.MG_TUNE_STEM_FUSED_STATS, int(os.environ.get("MGCONV_STEM_FUSED", "0")))
    gs, modes = [], []
    for c, m in segs:
        h = H // 2 if m == "u" else H
        g = Grid(ffi.MG_BF16, N, c, h, h)
        g.t.normal_()
        gs.append(g); modes.append(MG_SEG_UP if m == "u" else MG_SEG_SAME)
    pad = SHAPES[name][5] if len(SHAPES[name]) > 5 else (0 if k == 1 else 1)
    d = conv_desc(gs, modes, k, stride, pad, Cout, H, H)
    d.algo_fwd = d.algo_bwd_data = int(os.environ.get("MGCONV_ALGO", "0"))
    Ho = (H + 2 * pad - k) // stride + 1
    cin = sum(c for c, _ in segs)
    w = torch.randn(Cout, cin, k, k, device="cuda") * 0.05
    b = torch.zeros(Cout, device="cuda")
    wp = torch.zeros(ffi.lib.mg_conv_packed_bytes(C.byref(d), 0), dtype=torch.uint8, device="cuda")
    wpt = torch.zeros(max(16, ffi.lib.mg_conv_packed_bytes(C.byref(d), 1)), dtype=torch.uint8, device="cuda")
    ctx.call("mg_conv_pack_weights", C.byref(d), ptr(w), ptr(wp), 0)
    if stride == 1:
        ctx.call("mg_conv_pack_weights", C.byref(d), ptr(w), ptr(wpt), 1)
    y = Grid(ffi.MG_BF16, N, Cout, Ho, Ho)
    g = Grid(ffi.MG_BF16, N, Cout, Ho, Ho); g.t.normal_()
    cp = sum(x.Cp for x in gs)
    dcat = Grid(ffi.MG_BF16, N, cp, H, H, Cp=cp)
    dw = torch.zeros_like(w); db = torch.zeros_like(b)
    sums = torch.zeros(4 * Cout, dtype=torch.int64, device="cuda")   # 2 * Cout mg_sum
    flops = 2.0 * N * Ho * Ho * Cout * cin * k * k
    calls = {
        "fwd": lambda: ctx.call("mg_conv_forward", C.byref(d), ptr(w), ptr(wp), ptr(b), C.byref(y.g()), None),
        "fwd_stats": lambda: ctx.call("mg_conv_forward", C.byref(d), ptr(w), ptr(wp), ptr(b), C.byref(y.g()), ptr(sums)),
        "dgrad": lambda: ctx.call("mg_conv_backward_data", C.byref(d), ptr(w), ptr(wpt), C.byref(g.g()), C.byref(dcat.g())),
        "wgrad": lambda: ctx.call("mg_conv_backward_weight", C.byref(d), C.byref(g.g()), ptr(dw), None, 1.0),
    }
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for kind in which:
        f = calls[kind]
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush.zero_()  # evict L2 between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        out[kind] = {"us": round(ms * 1e3, 1), "tflops": round(flops / ms / 1e9, 1)}
    print(name, json.dumps(out), flush=True)


if __name__ == "__main__":
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else list(SHAPES)
    which = tuple(sys.argv[2].split(",")) if len(sys.argv) > 2 else ("fwd", "dgrad", "wgrad")
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    for n in names:
        run(n, which=which, iters=iters)
