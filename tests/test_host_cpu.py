"""CPU tests (-m "not gpu"): the C-ABI library loads and exports every symbol include/mgconv.h
declares, the host-side lowering reproduces the reference's published structure, and the
data-parallel host logic (sharding, bucket planning, parameter broadcast) works at world_size 2
on gloo.  No compute calls into the library are made here -- there is no GPU."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mgconv.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mgconv import ffi
    lib = ctypes.CDLL(ffi.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"libmgconv.so does not export {n}"
    assert set(names) == set(ffi.SIGNATURES), set(names) ^ set(ffi.SIGNATURES)
    assert lib.mg_version() >= 100


def test_header_parses_in_ffi_cdef_syntax():
    """lua/mgconv_ffi.lua feeds include/mgconv.h (minus preprocessor lines) to LuaJIT's ffi.cdef;
    Python cffi's ABI mode accepts the same C-declaration syntax: parse it and dlopen the library"""
    import cffi
    ffi = cffi.FFI()
    lines = [l for l in open(HEADER).read().split("\n")
             if not l.strip().startswith("#") and not l.startswith('extern "C"') and l.strip() != "}"]
    ffi.cdef("\n".join(lines).replace("size_t", "unsigned long"))
    lib = ffi.dlopen(os.path.join(ROOT, "multigrid-neural-architectures_b200", "mgconv", "libmgconv.so"))
    assert lib.mg_version() >= 100
    assert ffi.sizeof("mg_grid") == 48 and ffi.sizeof("mg_grad_src") == 64 and lib.MG_MAX_SEG == 6


def test_ctx_create_fails_loudly_without_gpu():
    from mgconv import ffi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ffi.MGError, match="no CPU fallback"):
        ffi.Context(0, 0, ffi.MG_BF16)


def test_struct_layout_matches_header():
    from mgconv import ffi
    assert ctypes.sizeof(ffi.mg_grid) == 48
    assert ctypes.sizeof(ffi.mg_conv_desc) == 8 + 6 * 48 + 6 * 4 + 6 * 4 + 2 * 4
    assert ctypes.sizeof(ffi.mg_grad_src) == 64


KNOWN = [  # netType, opt, input, params, conv+linear MACs  (README.md:85-92,109)
    ("cifar/nmg", dict(nLayer=1), (1, 3, 32, 32), 3_361_980, 51_707_520 + 560 * 100),
    ("cifar/rnmg", dict(nLayer=2), (1, 3, 32, 32), 17_524_920, 397_059_840 + 560 * 100),
    ("cifar/prnmg", dict(nLayer=2), (1, 3, 32, 32), 44_882_244, 1_031_777_280 + 896 * 100),
    ("ilsvrc/rnmg", dict(depth=34), (1, 3, 224, 224), 32_899_176, 5_759_765_760 + 512 * 1000),
]


@pytest.mark.parametrize("case", KNOWN, ids=[k[0] for k in KNOWN])
def test_builders_reproduce_published_structure(case):
    from mgconv import builders as B, lower as L
    name, opt, shape, params, macs = case
    m = B.load_net(name).createModel(B.Opt(nGPU=1, **opt))
    assert sum(w.numel() for w in m.parameters()[0]) == params
    b, _ = L.trace_model(m, shape)
    s = L.plan_summary(b)
    assert s["macs"] == macs
    assert s["max_segs"] <= 3


def test_ilsvrc_stage_shapes_and_fusion():
    """(224,112,56)->(56,28,14) ; ->(28,14,7) ; ->(14,7) ; ->(7)  models/ilsvrc/rnmg.lua:241,251-254;
    no pooled / up-sampled / concatenated tensor is materialised for a convolution"""
    from mgconv import builders as B, lower as L, ops as O
    from mgconv.ffi import MG_SEG_UP
    m = B.ilsvrc_rnmg.createModel(B.Opt(depth=34, nGPU=1))
    b, out = L.trace_model(m, (2, 3, 224, 224))
    convs = [o for o in b.ops if isinstance(o, O.ConvOp)]
    assert [(c.Cout, c.Ho) for c in convs[:3]] == [(64, 112), (32, 56), (16, 28)]
    assert (convs[3].Cout, convs[3].Ho, sum(t.C for t, _ in convs[3].segs)) == (64, 56, 96)
    assert (convs[-2].Cout, convs[-2].Ho) == (512, 7) and convs[-1].Cout == 1000
    # a 3-grid stage reads [pooled companion of the finer grid | same | UP of the coarser grid] ...
    c = convs[7]
    assert [(t.name, t.H, m) for t, m in c.segs] == [("conv3.act.pool", 28, 0), ("conv4.act", 28, 0), ("conv5.act", 14, MG_SEG_UP)]
    # ... and every gather is a pure copy: SAME or UP segments only, nothing pooled on the fly
    assert all(m in (0, MG_SEG_UP) for cv in convs for _, m in cv.segs)
    # the only JoinTable materialisations are the two shortcut operands after isConcat mgPools
    assert sum(isinstance(o, O.CatOp) for o in b.ops) == 2


def test_unmg_concat_unet_folds_into_segments():
    """ConcatUnet + MapTable(JoinTable(2)) (unmg.lua:219-220) never materialise a concatenation: the joined
    {shortcut_i, subnet_i} pairs, their pooled and up-sampled views all become segments (<= 6) of one conv"""
    from mgconv import builders as B, lower as L, ops as O
    m = B.load_net("mnist-cluttered/unmg").createModel(B.Opt(dataset="mnist-seg"))
    assert sum(w.numel() for w in m.parameters()[0]) == 5_903_290     # = the oracle's restatement of unmg.lua
    b, _ = L.trace_model(m, (2, 1, 64, 64))
    s = L.plan_summary(b)
    assert s["max_segs"] == 6 and s["convs"] == 20
    assert sum(isinstance(o, O.UpConvOp) for o in b.ops) == 6
    assert not any(isinstance(o, O.CatOp) for o in b.ops)


def test_shared_conv_output_is_detected():
    from mgconv import nn, lower as L
    conv = nn.SpatialConvolution(3, 4, 3, 3, 1, 1, 1, 1)
    m = nn.Sequential().add(conv).add(nn.ConcatTable().add(nn.ReLU(True)).add(nn.SpatialBatchNormalization(4)))
    with pytest.raises(NotImplementedError, match="two different consumers"):
        L.trace_model(m, (1, 3, 8, 8))


def test_concat_unet_zip():
    """layers/ConcatUnet.lua:7-37 incl. the #subnet < #shortcut case"""
    from mgconv import nn
    cu = nn.ConcatUnet()
    assert cu.trace([["t1", "t2", "t3"], ["p1", "p2"]], None) == [["t1", "p1"], ["t2", "p2"], ["t3"]]
    with pytest.raises(AssertionError):
        cu.trace([["t1"], ["p1", "p2"]], None)


def test_mgpool_mutates_its_argument_like_the_reference():
    from mgconv import builders as B
    n = [128, 64, 32]
    B.mgPool(n, True)
    assert n == [128, 96]   # models/ilsvrc/rnmg.lua:210-211
    n = [64, 32, 16]
    B.mgPool(n, False)
    assert n == [64, 32, 16]


def test_shard_and_bucket_planning():
    from mgconv.multigpu import shard_range, plan_buckets
    assert [shard_range(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]  # ceil(B/nGPU) chunks
    assert [shard_range(256, 8, r)[1] - shard_range(256, 8, r)[0] for r in range(8)] == [32] * 8
    sizes = [100, 50, 300, 20, 30, 500]
    bk = plan_buckets(sizes, bucket_bytes=4 * 300)
    cover = np.zeros(sum(sizes), dtype=int)
    for off, cnt, first in bk:
        cover[off:off + cnt] += 1
        assert off == sum(sizes[:first])
    assert (cover == 1).all()
    assert bk[0][0] + bk[0][1] == sum(sizes)          # first launched bucket ends the vector (last layers)
    assert [b[2] for b in bk] == sorted((b[2] for b in bk), reverse=True)


WORKER = r'''
import os, sys, torch, numpy as np
import torch.distributed as dist
sys.path.insert(0, os.path.join(sys.argv[1], "multigrid-neural-architectures_b200"))
from mgconv import builders as B, multigpu
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
torch.manual_seed(100 + rank)               # different initial weights per rank
m = B.cifar_rnmg.createModel(B.Opt(nLayer=1, nGPU=2))
assert isinstance(m, multigpu.DataParallel)
flat, gflat = m.getParameters()             # broadcast from rank 0 (syncParameters)
ref = flat.clone(); dist.broadcast(ref, 0)
assert torch.equal(flat, ref), "parameters differ after syncParameters"
# bucket trigger logic: replay a backward (reverse module order) against a recording engine
launched = []
class Ctx:
    def call(self, name, p, cnt, dbl): launched.append((name, cnt))
class Eng: ctx = Ctx()
m.model._engine = Eng()
mods = [mod for mod in m.model.listModules() if mod.own_parameters()]
import random
for order in (list(reversed(mods)), random.Random(rank).sample(mods, len(mods))):   # backward order, then any order
    del launched[:]
    m._reset_buckets()
    for mod in order:
        m._param_done(mod)
        m._param_done(mod)                  # idempotent
    assert sum(c for _, c in launched) == flat.numel(), (sum(c for _, c in launched), flat.numel())
    assert len(launched) == len(m.buckets)
lo, hi = multigpu.shard_range(7, 2, rank)
assert (lo, hi) == ((0, 4) if rank == 0 else (4, 7))

# ---- gradient accumulation (-iterSize 2), unequal shards (4 + 3 images) and a plan rebuilt for another batch size ----
# fake engines: backward ADDS a rank- and call-dependent gradient into gflat (accGradParameters semantics) and the
# "communicator" all-reduces the pointed-to range over gloo, so the wrapper's bookkeeping is checked end to end
import ctypes
calls = []
class FakeCtx:
    def __init__(self, eng): self.eng = eng
    def call(self, name, *a):
        calls.append((name, id(self.eng)))
        if name == "mg_allreduce_launch":
            off = (a[0].value - gflat.data_ptr()) // 4
            dist.all_reduce(gflat[off:off + a[1]])
class FakeEng:
    def __init__(self, key):
        self.key, self.ctx, self.on_param_done, self.bn_sync = key, FakeCtx(self), None, 0
        class D: index = 0
        self.device = D()
    def set_bn_sync(self, n): self.bn_sync = n
class FakeComm:
    h = 1
    def call(self, *a): calls.append(("comm:" + a[0], 0))
m._comm = FakeComm()          # stands for the context that owns the NCCL communicator
engines = {}
def get_engine(inp):
    m.model._engine = engines.setdefault(tuple(inp.shape), FakeEng(tuple(inp.shape)))
    return m.model._engine
step = [0]
def model_backward(inp, go, scale=1.0):
    step[0] += 1
    gflat.add_(scale * (rank + 1) * step[0])      # local gradient of this micro-batch
    eng = get_engine(inp)
    for mod in reversed(mods): eng.on_param_done(mod)
m.model._get_engine = get_engine
m.model.backward = model_backward
m.model.forward = lambda inp: None
xa = torch.zeros(4 if rank == 0 else 3, 3, 8, 8)
m.zeroGradParameters()
for it in range(2):                               # iterSize = 2: gradients zeroed only before the first micro-batch
    m.forward(xa); m.backward(xa, None)
w = [4 / 7, 3 / 7]
want = sum(w[r] * (r + 1) * s for r in range(2) for s in (1, 2))
assert torch.allclose(gflat, torch.full_like(gflat, want), rtol=1e-6), (float(gflat[0]), want)
# the user may also zero the flat gradient directly: the wrapper must not rely on zeroGradParameters() having been called
gflat.zero_(); step[0] = 0
m.forward(xa); m.backward(xa, None)
assert torch.allclose(gflat, torch.full_like(gflat, sum(w[r] * (r + 1) for r in range(2))), rtol=1e-6)
# another batch size (the partial last batch of the test pass): a new engine borrows the SAME communicator
xb = torch.zeros(2, 3, 8, 8)
m.zeroGradParameters(); m.forward(xb); m.backward(xb, None)
shared = [c for c in calls if c[0] == "mg_comm_share"]
assert len(shared) == 2 and shared[0][1] != shared[1][1], shared
assert not [c for c in calls if c[0].startswith("comm:")], "the communicator must be initialised once, not per engine"
# an empty shard is an error on every rank, not a hang
try:
    m.forward(torch.zeros(1 if rank == 0 else 0, 3, 8, 8)); raise SystemExit("empty shard accepted")
except multigpu.ffi.MGError as e:
    assert "without images" in str(e)
# wrapper methods that return self keep the wrapper
assert m.clearState() is m
dist.barrier()
print("ok", rank)
'''


def test_data_parallel_host_logic_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"ok {r}" in o, o


def test_lane_schedule_orders_every_conflicting_pair():
    """sched.Schedule: random op graphs over shared buffers -> replay the emitted lane / event calls on a model of
    CUDA stream semantics and check that every RAW / WAR / WAW pair is ordered, and that the pass ends joined on lane 0"""
    import random
    from mgconv import sched

    class FakeOp:
        def __init__(self, i, size, r, w):
            self.i, self._size, self.r, self.w = i, size, r, w

        def size(self):
            return self._size

    class FakeCtx:
        def __init__(self):
            self.lane, self.log = 0, []

        def call(self, name, *a):
            self.log.append((name, a))
            if name == "mg_ctx_lane":
                self.lane = a[0]

    class FakeE:
        pass

    rnd = random.Random(7)
    for trial in range(30):
        nbuf = rnd.randint(3, 10)
        ops = []
        for i in range(rnd.randint(5, 60)):
            r = rnd.sample(range(1, nbuf + 1), rnd.randint(0, min(3, nbuf)))
            w = rnd.sample(range(1, nbuf + 1), rnd.randint(1, 2))
            ops.append(FakeOp(i, rnd.choice([56, 28, 14, 7]), r, w))
        E = FakeE()
        E.ctx = FakeCtx()
        ran = []
        S = sched.Schedule(ops, lambda o: (o.r, o.w), lambda o: (lambda E_, o=o: ran.append((o.i, E_.ctx.lane, len(E_.ctx.log)))), 3, 100)
        S.run(E)
        assert [i for i, _, _ in ran] == list(range(len(ops)))          # plan order preserved
        assert E.ctx.lane == 0
        # model: done[lane] = set of ops known complete at the current point of that lane's stream
        done = {l: set() for l in range(4)}
        ev = {}
        pos = {i: (lane, at) for i, lane, at in ran}
        before = {}
        lane, k = 0, 0
        marks = sorted((at, i) for i, _, at in ran)
        mi = 0
        for n, (name, a) in enumerate(E.ctx.log + [("end", ())]):
            while mi < len(marks) and marks[mi][0] == n:     # op issued at this log position on the current lane
                i = marks[mi][1]
                before[i] = set(done[lane])
                done[lane].add(i)
                mi += 1
            if name == "mg_ctx_lane":
                lane = a[0]
            elif name == "mg_ctx_event_record":
                ev[a[0]] = set(done[lane])
            elif name == "mg_ctx_event_wait":
                done[lane] |= ev[a[0]]
        for j, oj in enumerate(ops):
            for i in range(j):
                oi = ops[i]
                conflict = (set(oi.w) & (set(oj.r) | set(oj.w))) or (set(oi.r) & set(oj.w))
                if conflict:
                    assert i in before[j], (trial, i, j)
        assert done[0] >= set(range(len(ops)))                              # final join


def test_modelfuncs_initialisers_statistics():
    """utils/modelfuncs.lua:3-54: MSRinit(model, backend, pow) is the FAN-IN normal with an exponent, XAVinit the uniform
    const * sqrt(sqrt / (nIn + nOut)), GAUSSinit N(mean, stddev); every one zeroes the conv biases; BNinit / FCinit /
    DisableBias as the Lua does.  Statistics are checked on the layers with enough weights for a 5 % bar."""
    import math
    import torch
    from mgconv import modelfuncs as MF, builders as B
    torch.manual_seed(0)
    model = B.load_net("cifar/rnmg").createModel(B.Opt(nGPU=1, nLayer=1))
    convs = model.findModules("cudnn.SpatialConvolution")
    assert len(convs) > 10
    big = [v for v in convs if v.weight.numel() >= 50000]
    for pw in (0.5, 0.3):
        MF.MSRinit(model, "cudnn", pw)
        for v in big:
            want = math.pow(2.0 / (v.kW * v.kH * v.nInputPlane), pw)
            assert abs(float(v.weight.std()) / want - 1) < 0.05 and abs(float(v.weight.mean())) < 0.05 * want
        assert all(float(v.bias.abs().max()) == 0 for v in convs)
    MF.MSRinit(model, "nn", 0.5)      # either backend spelling reaches the same modules (cudnn.convert leaves both in the wild)
    MF.XAVinit(model, "cudnn", 1.0, 6.0)
    for v in big:
        val = math.sqrt(6.0 / (v.nInputPlane + v.nOutputPlane))
        assert float(v.weight.abs().max()) <= val and abs(float(v.weight.std()) / (val / math.sqrt(3)) - 1) < 0.05
    MF.GAUSSinit(model, "cudnn", 0.25, 0.02)
    for v in big:
        assert abs(float(v.weight.mean()) - 0.25) < 2e-3 and abs(float(v.weight.std()) / 0.02 - 1) < 0.05
    MF.BNinit(model, "nn", 0.5, 0.125)
    bns = model.findModules("nn.SpatialBatchNormalization")
    assert bns and all(float(m.weight.min()) == 0.5 == float(m.weight.max()) and float(m.bias.min()) == 0.125 for m in bns)
    for m in model.findModules("nn.Linear"):
        m.bias.fill_(3.0)
    MF.FCinit(model)
    assert all(float(m.bias.abs().max()) == 0 for m in model.findModules("nn.Linear"))
    convs[0].bias.fill_(1.0)
    MF.DisableBias(model, "cudnn")
    assert float(convs[0].bias.abs().max()) == 0 and convs[0].noBias
    with pytest.raises(AssertionError):
        MF.MSRinit(model, "cunn", 0.5)    # modelfuncs.lua:4 asserts the backend name


def test_put2gpu_mirrors_utilfuncs_structure():
    """utils/utilfuncs.lua:19-30: a tensor destination takes exactly one CPU tensor, anything else is the Lua's error"""
    import torch
    from mgconv import utilfuncs as U
    with pytest.raises(Exception):
        U.put2GPU([torch.zeros(2), torch.zeros(2)], torch.zeros(2))
