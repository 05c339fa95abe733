"""Torch7 `.t7` interchange (mgconv/t7.py): known-answer byte strings assembled by hand from the
wire format of torch7's File.lua, round trips, and checkpoints of a whole multigrid model (the format saveDataParallel /
loadDataParallel use, multigpu.lua:105-160).  No Torch7 is available here, so the byte-level vectors are the anchor."""
import struct

import numpy as np
import pytest
import torch

from mgconv import builders as B, t7

i32 = lambda v: struct.pack("<i", v)
i64 = lambda v: struct.pack("<q", v)
f64 = lambda v: struct.pack("<d", v)
s = lambda x: i32(len(x)) + x.encode()


def test_scalars_and_strings_known_bytes():
    assert t7.dumps(None) == i32(0) and t7.loads(i32(0)) is None
    assert t7.dumps(3) == i32(1) + f64(3.0) and t7.loads(i32(1) + f64(3.0)) == 3 and isinstance(t7.loads(i32(1) + f64(3.0)), int)
    assert t7.loads(i32(1) + f64(0.1)) == 0.1
    assert t7.dumps(True) == i32(5) + i32(1) and t7.loads(i32(5) + i32(0)) is False
    assert t7.dumps("nn.ReLU") == i32(2) + s("nn.ReLU") and t7.loads(i32(2) + s("abc")) == "abc"


def test_table_and_shared_reference_known_bytes():
    # { [1] = "a", x = {}, y = <the same table as x> }: the second occurrence is written as its index only
    inner = i32(3) + i32(2) + i32(0)
    blob = (i32(3) + i32(1) + i32(3) +
            i32(1) + f64(1.0) + i32(2) + s("a") +
            i32(2) + s("x") + inner +
            i32(2) + s("y") + i32(3) + i32(2))
    v = t7.loads(blob)
    assert v[1] == "a" and v["x"] is v["y"] and v.array() == ["a"]
    shared = t7.T7Table()
    assert t7.dumps(t7.T7Table({1: "a", "x": shared, "y": shared})) == blob


def test_float_tensor_known_bytes_and_views():
    data = np.arange(6, dtype=np.float32)
    storage = i32(4) + i32(2) + s("V 1") + s("torch.FloatStorage") + i64(6) + data.tobytes()
    tensor = i32(4) + i32(1) + s("V 1") + s("torch.FloatTensor") + i32(2) + i64(2) + i64(3) + i64(3) + i64(1) + i64(1) + storage
    a = t7.loads(tensor)
    assert a.dtype == np.float32 and np.array_equal(a, data.reshape(2, 3))
    assert t7.dumps(data.reshape(2, 3)) == tensor
    # a transposed view with a storage offset: size (3, 2), stride (1, 3) ... and offset 2 (1-based) into an 8-element storage
    st8 = i32(4) + i32(2) + s("V 1") + s("torch.DoubleStorage") + i64(8) + np.arange(8, dtype=np.float64).tobytes()
    view = i32(4) + i32(1) + s("V 1") + s("torch.DoubleTensor") + i32(2) + i64(3) + i64(2) + i64(1) + i64(3) + i64(2) + st8
    assert np.array_equal(t7.loads(view), np.arange(8.0)[1:7].reshape(2, 3).T)
    # CudaTensor = float32 payload; class name without a version header (files older than "V 1")
    legacy = i32(4) + i32(1) + s("torch.CudaTensor") + i32(1) + i64(2) + i64(1) + i64(1) + i32(4) + i32(2) + s("torch.CudaStorage") + i64(2) + np.float32([7, 8]).tobytes()
    assert np.array_equal(t7.loads(legacy), np.float32([7, 8]))
    # empty tensor: nil storage
    empty = i32(4) + i32(1) + s("V 1") + s("torch.FloatTensor") + i32(0) + i64(1) + i32(0)
    assert t7.loads(empty).size == 0 and t7.dumps(np.zeros((0,), np.float32))[:len(empty) - 16] == empty[:len(empty) - 16]


def test_module_object_known_bytes_and_function_passthrough():
    fields = i32(3) + i32(2) + i32(2) + i32(2) + s("inplace") + i32(5) + i32(1) + i32(2) + s("train") + i32(5) + i32(0)
    blob = i32(4) + i32(1) + s("V 1") + s("nn.ReLU") + fields
    o = t7.loads(blob)
    assert o.typename == "nn.ReLU" and o.fields == {"inplace": True, "train": False}
    assert t7.dumps(o) == blob
    fn = i32(8) + i32(1) + i32(4) + b"\x1bLJ\x02" + i32(3) + i32(2) + i32(0)     # string.dump bytes + an empty upvalue table
    f = t7.loads(fn)
    assert isinstance(f, t7.T7Function) and f.dumped == b"\x1bLJ\x02" and t7.dumps(f) == fn


def test_round_trip_nested():
    r = np.random.default_rng(0)
    obj = {"epoch": 12, "lr": 0.05, "name": "rnmg", "flags": [True, False, None], "w": r.standard_normal((3, 4, 2)).astype(np.float32),
           "idx": np.arange(5, dtype=np.int64), "mod": t7.T7Object("nn.Identity", t7.T7Table({"train": True}))}
    back = t7.loads(t7.dumps(obj))
    assert back["epoch"] == 12 and back["lr"] == 0.05 and back["name"] == "rnmg"
    assert back["flags"].array() == [True, False] and 3 not in back["flags"]       # a nil value ends the array part, as in Lua
    assert np.array_equal(back["w"], obj["w"]) and np.array_equal(back["idx"], obj["idx"])
    assert back["mod"].typename == "nn.Identity" and back["mod"].fields["train"] is True


def test_errors_are_reported():
    good = t7.dumps({"a": np.ones(4, np.float32)})
    with pytest.raises(t7.T7Error, match="truncated"):
        t7.loads(good[:-3])
    with pytest.raises(t7.T7Error, match="trailing"):
        t7.loads(good + b"\0")
    with pytest.raises(t7.T7Error, match="unknown type tag"):
        t7.loads(i32(42))
    bad_view = i32(4) + i32(1) + s("V 1") + s("torch.FloatTensor") + i32(1) + i64(9) + i64(1) + i64(1) + i32(4) + i32(2) + s("V 1") + s("torch.FloatStorage") + i64(2) + np.float32([1, 2]).tobytes()
    with pytest.raises(t7.T7Error, match="exceeds its storage"):
        t7.loads(bad_view)


@pytest.mark.parametrize("net,opt", [("cifar/nmg", dict(nLayer=1)), ("cifar/rnmg", dict(nLayer=1)), ("ilsvrc/rnmg", dict(depth=34))])
def test_model_checkpoint_round_trip(tmp_path, net, opt):
    """save_model writes the module tree under the upstream class names; load_into restores every parameter and running
    statistic into a freshly built model, also through an nn.DataParallelTable wrapper (loadDataParallel, multigpu.lua:141-145)"""
    torch.manual_seed(1)
    N = B.load_net(net)
    a = N.createModel(B.Opt(nGPU=1, **opt))
    for m in a.listModules():
        for name in ("weight", "bias", "running_mean", "running_var"):
            t = getattr(m, name, None)
            if isinstance(t, torch.Tensor):
                t.uniform_(0.5, 1.5)
    path = str(tmp_path / "model_1.t7")
    t7.save_model(path, a)
    root = t7.load(path)
    assert root.typename == "nn.Sequential"
    convs = [o for o in t7._walk(root) if isinstance(o, t7.T7Object) and o.typename == "cudnn.SpatialConvolution"]
    assert len(convs) == len(a.findModules("cudnn.SpatialConvolution")) and convs[0].fields["weight"].ndim == 4 and convs[0].fields["groups"] == 1
    for wrap in (False, True):
        b = N.createModel(B.Opt(nGPU=1, **opt))
        src = t7.T7Object("nn.DataParallelTable", t7.T7Table({"modules": t7.T7Table({1: root}), "gpuAssignments": t7.T7Table({1: 1})})) if wrap else path
        t7.load_into(b, src)
        for ma, mb in zip(a.listModules(), b.listModules()):
            for name in ("weight", "bias", "running_mean", "running_var"):
                ta, tb = getattr(ma, name, None), getattr(mb, name, None)
                if isinstance(ta, torch.Tensor):
                    assert torch.equal(ta, tb), (ma.typename, name)
    # a checkpoint of a different architecture is rejected, not half-loaded
    other = B.load_net("cifar/nmg").createModel(B.Opt(nGPU=1, nLayer=2))
    with pytest.raises(t7.T7Error):
        t7.load_into(other, path)


@pytest.mark.parametrize("ext", ["pt", "t7"])
def test_checkpoint_after_a_forward_pass_round_trips(tmp_path, ext):
    """saveDataParallel after training steps: the top-level module then holds .output / .gradInput tensors, which the
    reference clears before saving (multigpu.lua:110-131); neither container may store them and both must load into a
    freshly built model (whose .output is None) -- native format and .t7"""
    from mgconv import multigpu
    torch.manual_seed(3)
    N = B.load_net("cifar/rnmg")
    opt = B.Opt(nGPU=1, nLayer=1)
    a = N.createModel(opt)
    for m in a.listModules():
        for name in ("weight", "bias", "running_mean", "running_var"):
            t = getattr(m, name, None)
            if isinstance(t, torch.Tensor):
                t.uniform_(0.5, 1.5)
    a.output, a.gradInput = torch.randn(4, 100), torch.randn(4, 3, 32, 32)      # state left behind by forward / backward
    path = str(tmp_path / ("model_3." + ext))
    multigpu.saveDataParallel(path, a)
    assert a.output is not None                                                  # saving does not disturb the live model
    b = multigpu.loadAndRemoveDPT(path, N, opt)
    assert b.output is None
    n = 0
    for ma, mb in zip(a.listModules(), b.listModules()):
        for name in ("weight", "bias", "running_mean", "running_var"):
            ta, tb = getattr(ma, name, None), getattr(mb, name, None)
            if isinstance(ta, torch.Tensor):
                assert torch.equal(ta, tb), (ma.typename, name)
                n += 1
    assert n > 20
    if ext == "pt":   # a file whose shapes do not fit is rejected with a clear error
        other = B.load_net("cifar/rnmg").createModel(B.Opt(nGPU=1, nLayer=2))
        with pytest.raises(Exception):
            multigpu._load_into(other, path)
