"""utils/utilfuncs.lua: put2GPU / recursivePut2Gpu (3-30) -- the host -> device staging of a batch -- plus the cheap parts of
the data hooks on the device (random crop, horizontal flip, mean / std normalisation; dataset/*/donkey.lua).

The reference resizes persistent CudaTensors and copies into them synchronously, once per trainBatch, after which it
synchronises the device (pipelines/standard/train.lua:124-135).  `Put2GPU` keeps that contract -- the caller gets persistent
device tensors holding the batch -- but stages through PINNED host buffers on a copy stream, double buffered, so that batch
i+1 crosses PCIe while step i trains; `wait()` makes the compute stream wait for the batch it is about to read.
"""
import ctypes as C

import torch

from . import ffi
from .ffi import ptr


def recursivePut2Gpu(Atable, AgpuTable):
    """utilfuncs.lua:3-17: nested tables of CPU tensors -> the same structure of persistent device tensors"""
    for i, a in enumerate(Atable):
        if isinstance(a, (list, tuple)):
            if i >= len(AgpuTable):
                AgpuTable.append([])
            recursivePut2Gpu(a, AgpuTable[i])
        else:
            if i >= len(AgpuTable):
                AgpuTable.append(torch.empty(0, device="cuda"))
            if tuple(AgpuTable[i].shape) != tuple(a.shape) or AgpuTable[i].dtype != a.dtype:
                AgpuTable[i] = torch.empty(a.shape, dtype=a.dtype, device=AgpuTable[i].device)
            AgpuTable[i].copy_(a)
    return AgpuTable


def put2GPU(cpuData, gpuLocation):
    """utilfuncs.lua:19-30: a list destination is filled recursively; a tensor destination takes the single CPU tensor"""
    if isinstance(gpuLocation, list):
        return recursivePut2Gpu(cpuData, gpuLocation)
    if len(cpuData) == 1:
        if tuple(gpuLocation.shape) != tuple(cpuData[0].shape):
            gpuLocation.resize_(cpuData[0].shape)
        gpuLocation.copy_(cpuData[0])
        return gpuLocation
    raise ffi.MGError("put2GPU: a tensor destination takes exactly one CPU tensor")   # 'Some kind of error...' (utilfuncs.lua:28)


class Put2GPU:
    """double-buffered put2GPU: stage(i, tensors...) enqueues the H2D copies of batch i on a copy stream into buffer set i & 1
    (waiting until the step that last read that set has finished); get(i) makes the current stream wait for them and returns the
    device tensors; done(i) marks the set free again.  Host tensors must be pinned for the copies to overlap the step."""

    def __init__(self, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.stream = torch.cuda.Stream(self.device)
        self.bufs = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]
        for e in self.consumed:
            e.record()
        self.bytes_per_batch = 0

    def stage(self, i, *host):
        b = i & 1
        if self.bufs[b] is None or any(tuple(d.shape) != tuple(h.shape) or d.dtype != h.dtype for d, h in zip(self.bufs[b], host)):
            self.bufs[b] = [torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host]
        self.bytes_per_batch = sum(h.numel() * h.element_size() for h in host)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.consumed[b])          # the step that last read this set is done
            for d, h in zip(self.bufs[b], host):
                d.copy_(h, non_blocking=True)
            self.ready[b].record(self.stream)

    def get(self, i):
        b = i & 1
        torch.cuda.current_stream(self.device).wait_event(self.ready[b])
        return self.bufs[b]

    def done(self, i):
        self.consumed[i & 1].record()


_aug_ctx = {}


def crop_flip_normalize(batch, out_hw, y0=None, x0=None, flip=None, mean=None, std=None, out=None):
    """the train / test hooks of dataset/*/donkey.lua on a device batch [N,C,H,W] (fp32): per-image crop window origin
    (y0, x0: int32 device tensors or None), horizontal flip flags, per-channel mean / std; positions outside the source image
    read as zero (the test hook's zero padding).  One kernel (mg_crop_flip_normalize), no host synchronisation."""
    if not batch.is_cuda:
        raise ffi.MGError("crop_flip_normalize works on the device batch put2GPU produced; there is no CPU fallback")
    N, Cc, H, W = batch.shape
    oH, oW = out_hw
    dev = batch.device.index
    if dev not in _aug_ctx:
        _aug_ctx[dev] = ffi.Context(dev, 0, ffi.MG_F32)
    ctx = _aug_ctx[dev]
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    if out is None:
        out = torch.empty((N, Cc, oH, oW), dtype=torch.float32, device=batch.device)
    f32 = lambda t: None if t is None else torch.as_tensor(t, dtype=torch.float32, device=batch.device).contiguous()
    i32 = lambda t: None if t is None else torch.as_tensor(t, dtype=torch.int32, device=batch.device).contiguous()
    y0, x0, flip, mean, std = i32(y0), i32(x0), i32(flip), f32(mean), f32(std)
    ctx.call("mg_crop_flip_normalize", ptr(batch.contiguous().float()), N, Cc, H, W, ptr(out), oH, oW, ptr(y0), ptr(x0), ptr(flip), ptr(mean), ptr(std))
    return out
