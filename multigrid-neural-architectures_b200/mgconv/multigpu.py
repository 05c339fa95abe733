"""Data parallelism: the four global helpers of multigpu.lua (makeDataParallel:81,
saveDataParallel:105, loadDataParallel:137, loadAndRemoveDPT:150) on a process-per-GPU design.

The reference wraps the model in nn.DataParallelTable(1, true, true): one process, one Lua
thread per GPU, batch split ceil(B/nGPU) along dim 1, nccl.reduce of the flat gradient to GPU 1
after backward, SGD on GPU 1, nccl.bcast of the parameters (pipelines/standard/train.lua:163-169).
Here every rank owns a full replica and its shard of the batch; the flat gradient is cut into
buckets in reverse layer order and each bucket is all-reduced (ncclAllReduce through libmgconv,
over NVLink) on a side stream as soon as the last wgrad that writes into it has been enqueued,
overlapping the rest of backward; every rank then applies the same SGD step, so no broadcast is
needed after initialisation.  BN statistics stay per replica, as in the reference (multigpu.lua:38-41).
"""
import ctypes as C
import os
import torch
import torch.distributed as dist

from . import ffi
from .ffi import ptr

BUCKET_BYTES = int(os.environ.get("MGCONV_BUCKET_MB", "16")) << 20


def shard_range(B, nGPU, rank):
    """rows of a global batch owned by `rank`: DataParallelTable splits dim 1 into ceil(B/nGPU) chunks"""
    per = -(-B // nGPU)
    lo = min(B, rank * per)
    return lo, min(B, lo + per)


def plan_buckets(sizes, bucket_bytes=BUCKET_BYTES, elt=4):
    """cut the flat vector (parameter sizes in module order) into contiguous buckets, built from the
    END of the vector because backward produces gradients in reverse module order.
    Returns [(offset, count, first_param_index)] in launch order; a bucket is complete when the
    gradient of parameter `first_param_index` (its lowest-indexed member) has been written."""
    buckets, end, count = [], sum(sizes), 0
    off = end
    for i in range(len(sizes) - 1, -1, -1):
        off -= sizes[i]
        count += sizes[i]
        if count * elt >= bucket_bytes or i == 0:
            buckets.append((off, count, i))
            count = 0
    return buckets


class DataParallel:
    """stands where nn.DataParallelTable stood: same module protocol, local shard in, local
    outputs out; gradients are global sums divided by nothing -- the criterion already averages
    over the *global* batch through gscale = 1/nranks (equal shards)."""
    typename = "nn.DataParallelTable"

    def __init__(self, model, nGPU, group=None, bnSync=False):
        self.model = model
        self.bnSync = bool(bnSync)
        self.nGPU = nGPU
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if self.world != nGPU:
            raise ffi.MGError(f"makeDataParallel: -nGPU {nGPU} but {self.world} ranks were launched "
                              "(one process per GPU: torchrun --nproc-per-node nGPU)")
        self.flat = self.gflat = None
        self.buckets = []
        self._comm_ready = False
        self._owner = {}
        self.needsSync = False
        self.train = True

    # ---- module protocol passthrough ----------------------------------------------------
    def __getattr__(self, name):
        return getattr(self.__dict__["model"], name)

    def listModules(self):
        return self.model.listModules()

    def training(self):
        self.train = True
        self.model.training()
        return self

    def evaluate(self):
        self.train = False
        self.model.evaluate()
        return self

    def getParameters(self):
        self.flat, self.gflat = self.model.getParameters()
        sizes, self._owner, idx = [], {}, 0
        for m in self.model.listModules():
            for name, w, _ in m.own_parameters():
                sizes.append(w.numel())
                self._owner.setdefault(id(m), []).append(idx)
                idx += 1
        self.sizes = sizes
        self.buckets = plan_buckets(sizes)
        self._trigger = {first: (off, cnt) for off, cnt, first in self.buckets}
        self.syncParameters()
        return self.flat, self.gflat

    def syncParameters(self):
        """nccl.bcast(root = GPU 1) of the flat parameter vector (train.lua:166-168): needed once,
        after initialisation / checkpoint load; afterwards every rank takes identical steps"""
        if self.world > 1 and self.flat is not None:
            if self.flat.is_cuda and dist.get_backend(self.group) == "gloo":
                tmp = self.flat.cpu()
                dist.broadcast(tmp, 0, group=self.group)
                self.flat.copy_(tmp)
            else:
                dist.broadcast(self.flat, 0, group=self.group)
        self.needsSync = False

    def _init_comm(self, eng):
        if self._comm_ready or self.world == 1:
            return
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            ffi_rc = ffi.lib.mg_comm_unique_id(C.c_void_p(uid.data_ptr()))
            if ffi_rc != 0:
                raise ffi.MGError("mg_comm_unique_id failed (libnccl.so.2 not loadable?)")
        if dist.get_backend(self.group) == "nccl":
            u = uid.cuda()
            dist.broadcast(u, 0, group=self.group)
            uid = u.cpu()
        else:
            dist.broadcast(uid, 0, group=self.group)
        eng.ctx.call("mg_comm_init", self.rank, self.world, C.c_void_p(uid.data_ptr()))
        self._comm_ready = True

    def forward(self, input):
        out = self.model.forward(input)
        eng = self.model._engine
        if self.world > 1 and eng.on_param_done is None and self.gflat is not None:
            self._init_comm(eng)
            eng.on_param_done = self._param_done
            if self.bnSync:
                # -bnSync: BatchNorm statistics over the GLOBAL batch (one tiny all-reduce of the per-channel
                # sums per BN layer, forward and backward) -- equals the single-device result on the whole
                # batch.  The first forward above still used local statistics: redo it.
                eng.bn_sync = self.world
                out = self.model.forward(input)
        return out

    def backward(self, input, gradOutput, scale=1.0):
        if self.gflat is not None:
            self._reset_buckets()
        gi = self.model.backward(input, gradOutput, scale / self.world)
        if self.world > 1 and self.gflat is not None:
            eng = self.model._engine
            for b, (off, cnt, _) in enumerate(self.buckets):   # parameters the plan never touched
                if b not in self.launched:
                    eng.ctx.call("mg_allreduce_launch", ptr(self.gflat[off:off + cnt]), cnt, 0)
            eng.ctx.call("mg_allreduce_wait")
        return gi

    def _param_done(self, mod):
        """called by the engine right after the wgrad / BN-backward of `mod` was enqueued: a bucket whose
        last missing gradient this was is all-reduced now, overlapping the rest of backward"""
        eng = self.model._engine
        for idx in self._owner.get(id(mod), ()):
            if idx in self._done:
                continue
            self._done.add(idx)
            b = self._bucket_of[idx]
            self._missing[b] -= 1
            if self._missing[b] == 0:
                off, cnt, _ = self.buckets[b]
                eng.ctx.call("mg_allreduce_launch", ptr(self.gflat[off:off + cnt]), cnt, 0)
                self.launched.append(b)

    def _reset_buckets(self):
        self._done = set()
        self.launched = []
        firsts = sorted((f, i) for i, (_, _, f) in enumerate(self.buckets))
        self._bucket_of = {}
        self._missing = [0] * len(self.buckets)
        for k, (f, b) in enumerate(firsts):
            end = firsts[k + 1][0] if k + 1 < len(firsts) else len(self.sizes)
            for idx in range(f, end):
                self._bucket_of[idx] = b
                self._missing[b] += 1


def makeDataParallel(model, nGPU, net=None, bnSync=False):
    """multigpu.lua:81-103"""
    if nGPU > 1:
        if not dist.is_initialized():
            raise ffi.MGError("makeDataParallel(nGPU > 1) needs torch.distributed to be initialised: launch one "
                              "process per GPU (python -m torch.distributed.run --nproc-per-node nGPU ...)")
        return DataParallel(model, nGPU, bnSync=bnSync)
    return model


def _state(model):
    m = model.model if isinstance(model, DataParallel) else model
    st = []
    for mod in m.listModules():
        d = {k: v.detach().cpu().clone() for k, v in vars(mod).items()
             if isinstance(v, torch.Tensor) and not k.startswith("grad") and not k.startswith("_")}
        st.append((mod.typename, d))
    return st


def saveDataParallel(filename, model):
    """multigpu.lua:105-135: persist replica 1 with its buffers cleared.  The module graph is
    rebuilt by the builder; what is stored are the per-module tensors in listModules() order."""
    if not dist.is_initialized() or dist.get_rank() == 0:
        if str(filename).endswith(".t7"):      # the reference's own container: torch.save of the module tree (mgconv/t7.py)
            from . import t7
            t7.save_model(filename, model)
        else:
            torch.save({"format": "mgconv-b200/1", "modules": _state(model)}, filename)


def _load_into(model, filename):
    if str(filename).endswith(".t7"):          # a Torch7 checkpoint (model_<epoch>.t7, possibly an nn.DataParallelTable)
        from . import t7
        return t7.load_into(model, filename)
    blob = torch.load(filename, map_location="cpu")
    mods = (model.model if isinstance(model, DataParallel) else model).listModules()
    if len(mods) != len(blob["modules"]):
        raise ffi.MGError(f"{filename}: {len(blob['modules'])} modules stored, model has {len(mods)}")
    for mod, (tn, d) in zip(mods, blob["modules"]):
        if tn != mod.typename:
            raise ffi.MGError(f"{filename}: module type {tn} does not match {mod.typename}")
        for k, v in d.items():
            getattr(mod, k).copy_(v)
    return model


def loadDataParallel(filename, nGPU, net, opt):
    """multigpu.lua:137-148: rebuild through the net's createModel, load, wrap for nGPU"""
    o = type(opt)(opt)
    o["nGPU"] = 1
    model = _load_into(net.createModel(o), filename)
    return makeDataParallel(model, nGPU, net)


def loadAndRemoveDPT(filename, net, opt):
    """multigpu.lua:150-160: load a checkpoint as a plain single-GPU module"""
    o = type(opt)(opt)
    o["nGPU"] = 1
    return _load_into(net.createModel(o), filename)
