"""The plan-level C ABI driven from plain C (tests/cabi_unit.c): one residual multigrid unit forward + backward through
mg_plan_create / mg_stage_forward / mg_stage_backward with no Python between the calls, compared with the oracle's
restatement of mgConv (models/ilsvrc/rnmg.lua:91-159).  This is the call sequence lua/mgconv_nn.lua makes through LuaJIT FFI."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import builders as OB
from util import rel_err, bf16_round, emulate_bf16_storage

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multigrid-neural-architectures_b200", "mgconv", "libmgconv.so")
BIN = os.path.join(ROOT, "tests", "cabi_unit")


def build_cabi_unit():
    """gcc only: the driver is C99 and sees nothing but include/mgconv.h and the CUDA runtime's C API"""
    src = os.path.join(ROOT, "tests", "cabi_unit.c")
    if os.path.exists(BIN) and os.path.getmtime(BIN) >= max(os.path.getmtime(src), os.path.getmtime(LIB)):
        return BIN
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["gcc", "-std=c99", "-O1", src, "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(cuda, "include"), "-o", BIN, LIB,
           "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath," + os.path.dirname(LIB), "-Wl,-rpath," + os.path.join(cuda, "lib64")]
    subprocess.run(cmd, check=True)
    return BIN


def test_cabi_unit_compiles_as_c99():
    """CPU check: the header is valid C (not only C++) and the driver links against every plan-level symbol"""
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    assert os.path.exists(build_cabi_unit())


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
@pytest.mark.parametrize("cin,cout,hs", [([16, 8, 8], [32, 16, 8], [16, 8, 4]), ([24, 24], [24, 24], [14, 7])], ids=["padded-shortcut-3grids", "identity-2grids"])
def test_residual_unit_through_the_c_abi(tmp_path, precision, tol, cin, cout, hs):
    binary = build_cabi_unit()
    rng = np.random.default_rng(17)
    torch.manual_seed(17)
    n, N = len(cin), 3
    om = OB.res_mgConv(list(cin), list(cout), [3] * n).double()
    convs = [m for m in om.modules() if isinstance(m, torch.nn.Conv2d)]
    bns = [m for m in om.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    assert len(convs) == 2 * n and len(bns) == 2 * n       # construction order: layer 1 grids, then layer 2 grids
    with torch.no_grad():
        for c in convs:
            if precision == "bf16":
                c.weight.copy_(c.weight.to(torch.bfloat16).to(c.weight.dtype))
            c.bias.normal_(0, 0.1)
        for b in bns:
            b.weight.uniform_(0.5, 1.5); b.bias.normal_(0, 0.2)
    if precision == "bf16":
        emulate_bf16_storage(om)
    xs = [bf16_round(rng.standard_normal((N, c, h, h))) for c, h in zip(cin, hs)]
    oxs = [torch.from_numpy(x).requires_grad_() for x in xs]
    oy = om(oxs)
    gos = [bf16_round(rng.standard_normal(tuple(y.shape))) for y in oy]
    torch.autograd.backward(oy, [torch.from_numpy(g) for g in gos])

    pad = lambda v: list(v) + [0] * (4 - len(v))
    hdr = np.array([n, N, 1] + pad(cin) + pad(cout) + pad(hs) + pad(hs) + pad([3] * n), dtype=np.int32)
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(hdr.tobytes())
        for x in xs:
            f.write(np.ascontiguousarray(x, dtype=np.float32).tobytes())
        for g in gos:
            f.write(np.ascontiguousarray(g, dtype=np.float32).tobytes())
        for c, b in zip(convs, bns):
            for t in (c.weight, c.bias, b.weight, b.bias, torch.zeros_like(b.running_mean), torch.ones_like(b.running_var)):
                f.write(t.detach().numpy().astype(np.float32).tobytes())
    r = subprocess.run([binary, precision, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "cabi_unit ok" in r.stdout, r.stdout + r.stderr
    if precision == "bf16":
        assert "0 tcgen05" not in r.stdout, "bf16 convolutions must run on the tensor-core kernels: " + r.stdout
    out = np.fromfile(tmp_path / "out.bin", dtype=np.float32)
    pos = 0

    def take(shape):
        nonlocal pos
        k = int(np.prod(shape))
        v = out[pos:pos + k].reshape(shape); pos += k
        return v
    for y in oy:
        assert rel_err(take(tuple(y.shape)), y.detach().numpy()) <= tol
    for x in oxs:
        assert rel_err(take(tuple(x.shape)), x.grad.numpy()) <= 2 * tol
    for c, b in zip(convs, bns):
        assert rel_err(take(tuple(c.weight.shape)), c.weight.grad.numpy()) <= 2 * tol
        gb = take(tuple(c.bias.shape))
        assert np.abs(gb).max() <= 50 * tol * max(1.0, float(c.weight.grad.abs().max()))     # ~0 under training-mode BatchNorm
        assert rel_err(take(tuple(b.weight.shape)), b.weight.grad.numpy()) <= 2 * tol
        assert rel_err(take(tuple(b.bias.shape)), b.bias.grad.numpy()) <= 2 * tol
        assert rel_err(take(tuple(b.running_mean.shape)), b.running_mean.numpy()) <= max(tol, 1e-3)
        assert rel_err(take(tuple(b.running_var.shape)), b.running_var.numpy()) <= max(tol, 1e-3)
    assert pos == out.size
