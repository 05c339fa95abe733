"""device time / achieved bandwidth of the HBM-bound passes on R-MG-34 block-1 tensors (B=256)"""
import ctypes as C, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multigrid-neural-architectures_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from mgconv import ffi
from mgconv.ffi import ptr, mg_grad_src, MG_SEG_SAME, MG_SEG_POOL, MG_SEG_UP
from util import Grid
ctx = ffi.Context(0, torch.cuda.current_stream().cuda_stream, ffi.MG_BF16)
N, Cc, H = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 56
def G(c, h, cp=None):
    g = Grid(ffi.MG_BF16, N, c, h, h, Cp=cp); g.t.normal_(); return g
y, res, out, pooled = G(Cc, H), G(Cc, H), G(Cc, H), G(Cc, H // 2)
scale = torch.rand(y.Cp, device="cuda"); shift = torch.rand(y.Cp, device="cuda")
y.scale, y.shift = scale, shift
sums = torch.zeros(2 * Cc, dtype=torch.float64, device="cuda")
dcat1, dcat2, dfin = G(Cc + Cc // 2, H, cp=Cc + Cc // 2), G(Cc + Cc // 2 + Cc // 4 if False else Cc + Cc // 2, H // 2, cp=Cc + Cc // 2), G(Cc, H)
D, Gg = G(Cc, H), G(Cc, H)
src = (mg_grad_src * 3)()
src[0].g, src[0].c_offset, src[0].mode = dcat1.g(), 0, MG_SEG_SAME
src[1].g, src[1].c_offset, src[1].mode = dfin.g(), 0, MG_SEG_SAME
src[2].g, src[2].c_offset, src[2].mode = dcat2.g(), 0, MG_SEG_POOL
coef = torch.rand(3 * y.Cp, device="cuda"); mean = torch.rand(y.Cp, device="cuda"); inv = torch.rand(y.Cp, device="cuda")
gam = torch.rand(Cc, device="cuda"); dg = torch.zeros(Cc, device="cuda"); db = torch.zeros(Cc, device="cuda"); cdb = torch.zeros(Cc, device="cuda")
yraw = G(Cc, H)
E = 2 * N * H * H * Cc   # bytes of one bf16 tensor
calls = {
  "bn_stats (1 read)": (lambda: ctx.call("mg_bn_stats", C.byref(yraw.g()), ptr(sums)), 1.0),
  "apply+pool (2 reads, 1.25 writes)": (lambda: ctx.call("mg_residual_forward", C.byref(y.g()), C.byref(res.g()), 1, C.byref(out.g()), C.byref(pooled.g())), 3.25),
  "combine 3 src + mask + sums (~4.9 reads, 1 write)": (lambda: ctx.call("mg_grad_combine", C.byref(out.g()), 1, C.byref(yraw.g()), 3, src, C.byref(D.g()), ptr(sums)), 5.9),
  "bn_backward (2 reads, 1 write)": (lambda: ctx.call("mg_bn_backward", C.byref(yraw.g()), C.byref(D.g()), C.byref(Gg.g()), ptr(sums), N * H * H, ptr(gam), ptr(mean), ptr(inv), ptr(dg), ptr(db), 1.0, ptr(coef), ptr(cdb)), 3.0),
}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, (f, passes) in calls.items():
    for _ in range(3): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); ms = ts[len(ts) // 2]
    print(f"{name:55s} {ms*1e3:8.1f} us   {passes * E / ms / 1e9:7.2f} TB/s algorithmic")
