"""Torch7 binary serialisation (`torch.save` / `torch.load`, default binary mode) -- SURVEY.md section 8 f-3.

The reference snapshots `model_<epoch>.t7` with torch.save(filename, cleanDPT(model)) (multigpu.lua:105-135,
pipelines/standard/pipeline.lua:7) and resumes / evaluates with torch.load (multigpu.lua:137-160).  This module
reads and writes that container so that reference checkpoints load into the modules of mgconv.nn and models
trained here can be handed back.  Pure host code: numpy for the payloads, no Torch7 needed.

Wire format (torch7 File.lua writeObject / readObject, little endian, "long" = 8 bytes):

    object   := int32 type, payload
    type     := 0 nil | 1 number (float64) | 2 string (int32 n, n bytes) | 3 table | 4 torch object | 5 boolean (int32)
                | 6 function | 7 legacy recursive function | 8 recursive function
    table    := int32 index, [ int32 n, n x (object key, object value) ]         payload only the first time an index occurs
    torch    := int32 index, [ string version ("V 1"), string className, class payload ]
                  torch.XTensor  : int32 nDim, nDim x int64 size, nDim x int64 stride, int64 storageOffset (1-based), object storage
                  torch.XStorage : int64 n, n raw elements
                  any other class (nn.*, cudnn.*): one object = the table of the instance's fields
    function := [int32 index,] int32 n, n bytes of string.dump, object upvalues          (kept opaque)

`load` returns numbers as float / int (integral doubles become int), tables as T7Table (a dict; `.array()` gives the
1..n part as a list), tensors as numpy arrays (views honouring size / stride / offset) and every other torch class as
T7Object(typename, fields).  `save` accepts the same vocabulary plus torch tensors and Python lists.
"""
import struct

import numpy as np

TYPE_NIL, TYPE_NUMBER, TYPE_STRING, TYPE_TABLE, TYPE_TORCH, TYPE_BOOLEAN = 0, 1, 2, 3, 4, 5
TYPE_FUNCTION, LEGACY_TYPE_RECUR_FUNCTION, TYPE_RECUR_FUNCTION = 6, 7, 8

_ELT = {"Float": np.float32, "Double": np.float64, "Half": np.float16, "Byte": np.uint8, "Char": np.int8,
        "Short": np.int16, "Int": np.int32, "Long": np.int64,
        "Cuda": np.float32, "CudaDouble": np.float64, "CudaHalf": np.float16, "CudaByte": np.uint8, "CudaChar": np.int8,
        "CudaShort": np.int16, "CudaInt": np.int32, "CudaLong": np.int64}
_NAME_OF = {np.dtype(np.float32): "Float", np.dtype(np.float64): "Double", np.dtype(np.float16): "Half", np.dtype(np.uint8): "Byte",
            np.dtype(np.int8): "Char", np.dtype(np.int16): "Short", np.dtype(np.int32): "Int", np.dtype(np.int64): "Long"}


class T7Error(ValueError):
    pass


class T7Table(dict):
    """a Lua table; integer keys 1..n are its array part"""

    def array(self):
        out, i = [], 1
        while i in self:
            out.append(self[i])
            i += 1
        return out


class T7Object:
    """instance of a torch class that is serialised as the table of its fields (nn.Module subclasses, ...)"""

    def __init__(self, typename, fields=None):
        self.typename = typename
        self.fields = fields if fields is not None else T7Table()

    def __repr__(self):
        return f"T7Object({self.typename}, {sorted(map(str, self.fields))})"


class T7Function:
    """a serialised Lua function: opaque (bytecode + upvalues), written back verbatim"""

    def __init__(self, kind, dumped, upvalues):
        self.kind, self.dumped, self.upvalues = kind, dumped, upvalues


# ---------------------------------------------------------------- reader ----------------------------
class _Reader:
    def __init__(self, data):
        self.b, self.p, self.memo = memoryview(data), 0, {}

    def _take(self, n):
        if self.p + n > len(self.b):
            raise T7Error(f"truncated file: wanted {n} bytes at offset {self.p} of {len(self.b)}")
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def int32(self):
        return struct.unpack("<i", self._take(4))[0]

    def int64(self):
        return struct.unpack("<q", self._take(8))[0]

    def string(self):
        n = self.int32()
        if n < 0:
            raise T7Error(f"negative string length at offset {self.p}")
        return bytes(self._take(n)).decode("latin-1")

    def obj(self):
        t = self.int32()
        if t == TYPE_NIL:
            return None
        if t == TYPE_NUMBER:
            v = struct.unpack("<d", self._take(8))[0]
            return int(v) if v == int(v) and abs(v) < 2 ** 53 else v
        if t == TYPE_STRING:
            return self.string()
        if t == TYPE_BOOLEAN:
            return self.int32() != 0
        if t == TYPE_TABLE:
            idx = self.int32()
            if idx in self.memo:
                return self.memo[idx]
            tab = T7Table()
            self.memo[idx] = tab
            for _ in range(self.int32()):
                k = self.obj()
                tab[k] = self.obj()
            return tab
        if t == TYPE_TORCH:
            idx = self.int32()
            if idx in self.memo:
                return self.memo[idx]
            version = self.string()
            cls = self.string() if version.startswith("V ") else version      # files older than "V 1" carry the class name only
            return self._torch(idx, cls)
        if t in (TYPE_FUNCTION, LEGACY_TYPE_RECUR_FUNCTION, TYPE_RECUR_FUNCTION):
            idx = self.int32() if t != TYPE_FUNCTION else None
            if idx is not None and idx in self.memo:
                return self.memo[idx]
            f = T7Function(t, bytes(self._take(self.int32())), None)
            if idx is not None:
                self.memo[idx] = f
            f.upvalues = self.obj()
            return f
        raise T7Error(f"unknown type tag {t} at offset {self.p - 4}")

    def _torch(self, idx, cls):
        if cls.startswith("torch.") and cls.endswith("Storage"):
            dt = _ELT.get(cls[6:-7])
            if dt is None:
                raise T7Error(f"unsupported storage class {cls}")
            n = self.int64()
            a = np.frombuffer(self._take(n * np.dtype(dt).itemsize), dtype=np.dtype(dt).newbyteorder("<"), count=n).astype(dt)
            self.memo[idx] = a
            return a
        if cls.startswith("torch.") and cls.endswith("Tensor"):
            dt = _ELT.get(cls[6:-6])
            if dt is None:
                raise T7Error(f"unsupported tensor class {cls}")
            nd = self.int32()
            size = [self.int64() for _ in range(nd)]
            stride = [self.int64() for _ in range(nd)]
            off = self.int64() - 1
            self.memo[idx] = None        # reserve the index: the storage below takes the next one
            st = self.obj()
            if st is None or nd == 0:
                a = np.zeros(size if nd else (0,), dtype=dt)
            else:
                item = st.dtype.itemsize
                need = off + sum((s - 1) * k for s, k in zip(size, stride)) + 1 if all(s > 0 for s in size) else 0
                if off < 0 or need > st.size:
                    raise T7Error(f"{cls}: view (offset {off}, size {size}, stride {stride}) exceeds its storage of {st.size} elements")
                a = np.lib.stride_tricks.as_strided(st[off:], shape=size, strides=[k * item for k in stride], writeable=False)
            self.memo[idx] = a
            return a
        o = T7Object(cls)
        self.memo[idx] = o
        fields = self.obj()
        if not isinstance(fields, dict):
            raise T7Error(f"{cls}: expected the table of its fields, found {type(fields).__name__}")
        o.fields = fields
        return o


def loads(data):
    r = _Reader(data)
    v = r.obj()
    if r.p != len(r.b):
        raise T7Error(f"{len(r.b) - r.p} trailing bytes after the root object")
    return v


def load(filename):
    with open(filename, "rb") as f:
        return loads(f.read())


# ---------------------------------------------------------------- writer ----------------------------
class _Writer:
    def __init__(self):
        self.out, self.memo, self.next, self.keep = [], {}, 1, []

    def int32(self, v):
        self.out.append(struct.pack("<i", v))

    def int64(self, v):
        self.out.append(struct.pack("<q", v))

    def string(self, s):
        b = s.encode("latin-1")
        self.int32(len(b))
        self.out.append(b)

    def _index(self, o):
        """(index, already written?) -- objects are identified by id(), like Lua references"""
        k = id(o)
        if k in self.memo:
            return self.memo[k], True
        self.memo[k] = self.next
        self.keep.append(o)
        self.next += 1
        return self.memo[k], False

    def obj(self, o):
        try:
            import torch
            if isinstance(o, torch.Tensor):
                o = o.detach().cpu().numpy() if o.dtype != torch.bfloat16 else o.detach().float().cpu().numpy()
        except ImportError:
            pass
        if o is None:
            return self.int32(TYPE_NIL)
        if isinstance(o, (bool, np.bool_)):
            self.int32(TYPE_BOOLEAN)
            return self.int32(1 if o else 0)
        if isinstance(o, (int, float, np.integer, np.floating)):
            self.int32(TYPE_NUMBER)
            return self.out.append(struct.pack("<d", float(o)))
        if isinstance(o, str):
            self.int32(TYPE_STRING)
            return self.string(o)
        if isinstance(o, (list, tuple)):
            o = T7Table({i + 1: v for i, v in enumerate(o)})
        if isinstance(o, dict):
            self.int32(TYPE_TABLE)
            idx, seen = self._index(o)
            self.int32(idx)
            if not seen:
                items = [(k, v) for k, v in o.items() if v is not None]      # a Lua table cannot hold nil
                self.int32(len(items))
                for k, v in items:
                    self.obj(k)
                    self.obj(v)
            return
        if isinstance(o, np.ndarray):
            name = _NAME_OF.get(o.dtype)
            if name is None:
                raise T7Error(f"no Torch7 tensor type for dtype {o.dtype}")
            self.int32(TYPE_TORCH)
            idx, seen = self._index(o)
            self.int32(idx)
            if seen:
                return
            a = np.ascontiguousarray(o)
            self.string("V 1")
            self.string(f"torch.{name}Tensor")
            self.int32(a.ndim)
            for s in a.shape:
                self.int64(s)
            for s in a.strides:
                self.int64(s // a.itemsize)
            self.int64(1)
            if a.ndim == 0 or a.size == 0:
                return self.int32(TYPE_NIL)
            self.int32(TYPE_TORCH)
            self.int32(self.next)
            self.next += 1
            self.string("V 1")
            self.string(f"torch.{name}Storage")
            self.int64(a.size)
            return self.out.append(a.astype(a.dtype.newbyteorder("<")).tobytes())
        if isinstance(o, T7Object):
            self.int32(TYPE_TORCH)
            idx, seen = self._index(o)
            self.int32(idx)
            if not seen:
                self.string("V 1")
                self.string(o.typename)
                self.obj(o.fields)
            return
        if isinstance(o, T7Function):
            self.int32(o.kind)
            if o.kind != TYPE_FUNCTION:
                idx, seen = self._index(o)
                self.int32(idx)
                if seen:
                    return
            self.int32(len(o.dumped))
            self.out.append(o.dumped)
            return self.obj(o.upvalues)
        raise T7Error(f"cannot serialise {type(o).__name__}")


def dumps(obj):
    w = _Writer()
    w.obj(obj)
    return b"".join(w.out)


def save(filename, obj):
    with open(filename, "wb") as f:
        f.write(dumps(obj))


# ---------------------------------------------------------------- nn.Module <-> .t7 -------------------
_ALIASES = {"nn.SpatialConvolution": "cudnn.SpatialConvolution", "nn.SpatialConvolutionMM": "cudnn.SpatialConvolution",
            "cudnn.SpatialBatchNormalization": "nn.SpatialBatchNormalization", "nn.SpatialFullConvolution": "cudnn.SpatialFullConvolution",
            "cudnn.ReLU": "nn.ReLU", "nn.SpatialAveragePooling": "cudnn.SpatialAveragePooling", "cudnn.SpatialMaxPooling": "nn.SpatialMaxPooling"}


def _canon(typename):
    return _ALIASES.get(typename, typename)


def _children(o):
    """sub-modules of a loaded container, in order (nn.Container keeps them in the `modules` array)"""
    mods = o.fields.get("modules") if isinstance(o, T7Object) else None
    return mods.array() if isinstance(mods, T7Table) else []


def _walk(o):
    yield o
    for c in _children(o):
        yield from _walk(c)


def unwrap_dpt(root):
    """nn.DataParallelTable -> its first replica (loadDataParallel / loadAndRemoveDPT, multigpu.lua:141-157)"""
    while isinstance(root, T7Object) and root.typename == "nn.DataParallelTable":
        kids = _children(root)
        if not kids:
            raise T7Error("nn.DataParallelTable without replicas")
        root = kids[0]
    return root


def load_into(model, source):
    """copy the parameters and BatchNorm running statistics of a Torch7 checkpoint (filename or loaded object) into `model`
    (mgconv.nn module tree built by the same builder).  Modules are matched in listModules() order -- the builders here mirror
    the Lua ones container for container -- and every match is checked by class and shape."""
    import torch
    root = unwrap_dpt(load(source) if isinstance(source, (str, bytes)) and not isinstance(source, T7Object) else source)
    if not isinstance(root, T7Object):
        raise T7Error(f"checkpoint root is {type(root).__name__}, expected an nn.Module")
    src = [o for o in _walk(root) if isinstance(o, T7Object) and isinstance(o.fields.get("weight"), np.ndarray)]
    inner = getattr(model, "model", model)          # DataParallel wrapper
    dst = [m for m in inner.listModules() if m.own_parameters()]
    if len(src) != len(dst):
        raise T7Error(f"checkpoint has {len(src)} parametrised modules, the model has {len(dst)}")
    for i, (o, m) in enumerate(zip(src, dst)):
        if _canon(o.typename) != _canon(m.typename):
            raise T7Error(f"module {i}: checkpoint {o.typename} vs model {m.typename}")
        for name in ("weight", "bias", "running_mean", "running_var"):
            t = getattr(m, name, None)
            a = o.fields.get(name)
            if t is None or a is None:
                continue
            if not isinstance(a, np.ndarray) or a.size != t.numel():
                raise T7Error(f"module {i} ({o.typename}) {name}: checkpoint shape {getattr(a, 'shape', None)} vs model {tuple(t.shape)}")
            with torch.no_grad():
                t.copy_(torch.from_numpy(np.array(a, dtype=np.float32)).reshape(t.shape))
    return model


_SCALARS = ("nInputPlane", "nOutputPlane", "kW", "kH", "dW", "dH", "padW", "padH", "eps", "momentum", "index", "dimension", "nInputDims",
            "dim", "pad", "nInputDim", "value", "scale_factor", "p", "inplace", "ceil_mode", "count_include_pad", "divide", "v2", "train")


def to_t7(module):
    """mgconv.nn module tree -> T7Object tree with the field names of the upstream classes (what saveDataParallel would have
    written after clearing output / gradInput, multigpu.lua:105-135); tensors as float32"""
    import torch
    inner = getattr(module, "model", module)
    f = T7Table()
    for k in _SCALARS:
        if hasattr(inner, k):
            v = getattr(inner, k)
            if isinstance(v, (bool, int, float)):
                f[k] = v
    for k, v in vars(inner).items():
        # output / gradInput are cleared before saving (multigpu.lua:110-131): activations are never stored
        if isinstance(v, torch.Tensor) and not k.startswith("_") and k not in ("output", "gradInput"):
            f[k] = v.detach().float().cpu().numpy()
    if inner.typename == "cudnn.SpatialConvolution":
        f["groups"] = 1
    if inner.typename == "nn.SpatialBatchNormalization":
        f["affine"], f["nDim"] = True, 4
    kids = inner.children()
    if kids:
        f["modules"] = T7Table({i + 1: to_t7(c) for i, c in enumerate(kids)})
    return T7Object(inner.typename, f)


def save_model(filename, model):
    """saveDataParallel (multigpu.lua:105-135) in the reference's own container format"""
    save(filename, to_t7(model))
