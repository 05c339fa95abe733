"""The reference's model-builder API on the B200-native backend.

Every function keeps the name, argument meaning and side effects of the file-local Lua helper
it stands for (cited per function); the NET tables at the bottom satisfy the contract of
models/basic_model.lua:1-91 and are selected by the same -netType strings
(opts.lua:46, model.lua:23).  The graph they build is made of mgconv.nn modules, which lower to
fused libmgconv calls -- see nn.py / lower.py.
"""
import math

from . import nn
from .multigpu import makeDataParallel

Convolution = nn.SpatialConvolution
ReLU = nn.ReLU
Max = nn.SpatialMaxPooling
Avg = nn.SpatialAveragePooling
SBatchNorm = nn.SpatialBatchNormalization
UpSample = nn.SpatialUpSamplingNearest


# ------------------------------------------------------------------------------------ units
def Shortcut(nIP, nOP, allow_conv=False):
    """models/ilsvrc/rnmg.lua:13-20; 1x1 conv + BN when channels shrink:
    models/mnist-cluttered/prnmg.mnist.lua:13-25"""
    if nOP > nIP:
        return nn.Padding(1, nOP - nIP, 3)
    if nIP > nOP:
        if not allow_conv:
            raise ValueError("Shortcut: nInputPlane > nOutputPlane")
        return nn.Sequential().add(Convolution(nIP, nOP, 1, 1, 1, 1, 0, 0)).add(SBatchNorm(nOP))
    return nn.Identity()


def ConvBNReLU(mod, nIP, nOP, kernel, eps=1e-5):
    """models/ilsvrc/rnmg.lua:22-30; eps 1e-3 in models/cifar/nmg.lua:18-29"""
    pad = 0 if kernel == 1 else 1
    mod.add(Convolution(nIP, nOP, kernel, kernel, 1, 1, pad, pad))
    mod.add(SBatchNorm(nOP, eps))
    mod.add(ReLU(True))
    return mod


def ConvBN(mod, nIP, nOP, kernel, eps=1e-5):
    """models/ilsvrc/rnmg.lua:32-39"""
    pad = 0 if kernel == 1 else 1
    mod.add(Convolution(nIP, nOP, kernel, kernel, 1, 1, pad, pad))
    mod.add(SBatchNorm(nOP, eps))
    return mod


def ResampleConcat(nIPs, isDrop=False):
    """models/ilsvrc/rnmg.lua:41-89: per grid JoinTable(2){ Max(finer), same, UpSample(coarser) };
    isDrop (models/mnist-cluttered/prnmg.mnist.lua:44-92) emits grids 1..n-1 only.
    Returns the module and the per-grid concatenated widths."""
    resample_concat = nn.ConcatTable()
    nOPs = []
    nGrids = len(nIPs) - 1 if isDrop else len(nIPs)
    for iG in range(1, nGrids + 1):
        grid = nn.Sequential()
        multi_scales = nn.ConcatTable()
        nIP = 0
        if iG - 1 > 0:
            multi_scales.add(nn.Sequential().add(nn.SelectTable(iG - 1)).add(Max(2, 2, 2, 2, 0, 0).ceil()))
            nIP += nIPs[iG - 2]
        multi_scales.add(nn.SelectTable(iG))
        nIP += nIPs[iG - 1]
        if iG + 1 <= nGrids:
            multi_scales.add(nn.Sequential().add(nn.SelectTable(iG + 1)).add(UpSample(2)))
            nIP += nIPs[iG]
        grid.add(multi_scales)
        grid.add(nn.JoinTable(2))
        resample_concat.add(grid)
        nOPs.append(nIP)
    return resample_concat, nOPs


def mgConv_plain(nInputPlanes, nOutputPlanes, kernels, eps=1e-3, isReLU=True):
    """plain multigrid convolution: models/cifar/nmg.lua:31-86; unmg.lua:54-109 (isReLU)"""
    assert len(nInputPlanes) == len(nOutputPlanes) == len(kernels), "#nInputPlanes, #nOutputPlanes and #kernels differ"
    resample_concat, nIPs = ResampleConcat(nInputPlanes)
    multi_grids = nn.ConcatTable()
    for iG in range(len(nInputPlanes)):
        grid = nn.Sequential()
        grid.add(resample_concat.get(iG + 1))
        (ConvBNReLU if isReLU else ConvBN)(grid, nIPs[iG], nOutputPlanes[iG], kernels[iG], eps)
        multi_grids.add(grid)
    return multi_grids


def mgConv(nInputPlanes, nOutputPlanes, kernels, isDrop=False, isOut=False, conv_shortcut=False):
    """residual multigrid unit: models/ilsvrc/rnmg.lua:91-159 (= cifar/rnmg.lua:102-173,
    cifar/prnmg.lua:122-193); isDrop / isOut: models/mnist-cluttered/prnmg.mnist.lua:108-175"""
    if not isDrop:
        assert len(nInputPlanes) == len(nOutputPlanes), "#nInputPlanes is not consistent with #nOutputPlanes"
    mg_conv = nn.Sequential()
    shortcut_convs = nn.ConcatTable()

    convs = nn.Sequential()
    resample_concat, _nIPs = ResampleConcat(nInputPlanes, isDrop)
    convs.add(resample_concat)
    conv_bn_relu = nn.ParallelTable()
    for i in range(len(_nIPs)):
        conv_bn_relu.add(ConvBNReLU(nn.Sequential(), _nIPs[i], nOutputPlanes[i], kernels[i]))
    convs.add(conv_bn_relu)
    resample_concat, _nIPs = ResampleConcat(nOutputPlanes[:len(_nIPs)], False)
    convs.add(resample_concat)
    conv_bn = nn.ParallelTable()
    for i in range(len(_nIPs)):
        conv_bn.add(ConvBN(nn.Sequential(), _nIPs[i], nOutputPlanes[i], kernels[i]))
    convs.add(conv_bn)
    shortcut_convs.add(convs)

    nShortcut = len(_nIPs)
    shortcut = nn.ConcatTable()
    for i in range(nShortcut):
        shortcut.add(nn.Sequential().add(nn.SelectTable(i + 1))
                     .add(Shortcut(nInputPlanes[i], nOutputPlanes[i], conv_shortcut)))
    shortcut_convs.add(shortcut)

    add_shortcut_convs = nn.ConcatTable()
    for i in range(nShortcut):
        pick = nn.ConcatTable()
        pick.add(nn.Sequential().add(nn.SelectTable(1)).add(nn.SelectTable(i + 1)))
        pick.add(nn.Sequential().add(nn.SelectTable(2)).add(nn.SelectTable(i + 1)))
        s = nn.Sequential().add(pick).add(nn.CAddTable(True))
        if not isOut:
            s.add(ReLU(True))
        add_shortcut_convs.add(s)
    mg_conv.add(shortcut_convs)
    mg_conv.add(add_shortcut_convs)
    return mg_conv


def resConv(nIP, nOP, kernel, conv_shortcut=False):
    """single-grid residual pair: models/cifar/prnmg.lua:102-120; prnmg.mnist.lua:94-106"""
    s = nn.Sequential()
    ConvBNReLU(s, nIP, nOP, kernel)
    ConvBN(s, nOP, nOP, kernel)
    return (nn.Sequential()
            .add(nn.ConcatTable().add(s).add(Shortcut(nIP, nOP, conv_shortcut)))
            .add(nn.CAddTable(True))
            .add(ReLU(True)))


def mgPool(nInputPlanes, isConcat):
    """models/ilsvrc/rnmg.lua:191-224.  MUTATES nInputPlanes (210-211): with isConcat the pooled
    grid n-1 is joined with grid n and the pyramid loses its coarsest resolution."""
    mg_pool = nn.ConcatTable()
    nGrids = len(nInputPlanes)
    for i in range(1, nGrids + 1):
        proc = nn.Sequential()
        if i == nGrids - 1 and isConcat:
            pool_cat = nn.ConcatTable()
            pool_cat.add(nn.Sequential().add(nn.SelectTable(i)).add(Max(2, 2, 2, 2, 0, 0).ceil()))
            pool_cat.add(nn.SelectTable(i + 1))
            proc.add(pool_cat)
            proc.add(nn.JoinTable(2))
            nInputPlanes[i - 1] = nInputPlanes[i - 1] + nInputPlanes[i]
            del nInputPlanes[i]
            mg_pool.add(proc)
            break
        proc.add(nn.SelectTable(i))
        proc.add(Max(2, 2, 2, 2, 0, 0).ceil())
        mg_pool.add(proc)
    return mg_pool


def mgConvInput_pyramid(nOutputPlanes, nIn=3, eps=1e-5):
    """CIFAR / MNIST input: [AvgPool r] -> 3x3 conv -> BN -> ReLU per grid
    (models/cifar/nmg.lua:88-106 with eps 1e-3; cifar/prnmg.lua:195-213; prnmg.mnist.lua:177-195)"""
    mg_inputs = nn.ConcatTable()
    for iG in range(1, len(nOutputPlanes) + 1):
        proc = nn.Sequential()
        if iG == 1:
            proc.add(nn.Identity())
        else:
            r = 2 ** (iG - 1)
            proc.add(Avg(r, r, r, r, 0, 0))
        ConvBNReLU(proc, nIn, nOutputPlanes[iG - 1], 3, eps)
        mg_inputs.add(proc)
    return mg_inputs


def mgConvInput_ilsvrc(nOutputPlanes):
    """models/ilsvrc/rnmg.lua:161-189: [AvgPool r] -> 7x7 s2 p3 conv -> BN -> ReLU -> MaxPool 3x3 s2 p1"""
    resample_image = nn.ConcatTable()
    for i in range(1, len(nOutputPlanes) + 1):
        proc = nn.Sequential()
        if i == 1:
            proc.add(nn.Identity())
        else:
            r = 2 ** (i - 1)
            proc.add(Avg(r, r, r, r, 0, 0))
        nOP = nOutputPlanes[i - 1]
        proc.add(Convolution(3, nOP, 7, 7, 2, 2, 3, 3))
        proc.add(SBatchNorm(nOP))
        proc.add(ReLU(True))
        proc.add(Max(3, 3, 2, 2, 1, 1))
        resample_image.add(proc)
    return nn.Sequential().add(resample_image)


def mgConvInput_cifar_rnmg(nOutputPlanes):
    """models/cifar/rnmg.lua:175-254: pyramid convs followed by one residual unit"""
    m = nn.Sequential()
    m.add(mgConvInput_pyramid(nOutputPlanes, 3))
    unit = mgConv(nOutputPlanes, nOutputPlanes, [3] * len(nOutputPlanes))
    for sub in unit.modules:
        m.add(sub)
    return m


def MultiGridsInput(model, nOPs, nLayer, nIn, conv_shortcut=False):
    """progressive schedule (models/cifar/prnmg.lua:258-307; prnmg.mnist.lua:205-252): after the
    pyramid convs run nLayer residual convs on the coarsest grid, then nLayer residual mg units on
    the two coarsest, ... ; untouched finer grids pass through SelectTable + FlattenTable."""
    model.add(mgConvInput_pyramid(nOPs, nIn))
    n = len(nOPs)
    for nGrid in range(1, n + 1):
        for _ in range(nLayer):
            if nGrid > 1:
                mg_convs = nn.ConcatTable()
                for j in range(1, n - nGrid + 1):
                    mg_convs.add(nn.SelectTable(j))
                _select = nn.ConcatTable()
                _nOPs = []
                for j in range(n - nGrid + 1, n + 1):
                    _select.add(nn.SelectTable(j))
                    _nOPs.append(nOPs[j - 1])
                _mg_conv = nn.Sequential().add(_select)
                _mg_conv.add(mgConv(_nOPs, _nOPs, [3] * len(_nOPs), conv_shortcut=conv_shortcut))
                mg_convs.add(_mg_conv)
                model.add(mg_convs)
                model.add(nn.FlattenTable())
            else:
                convs = nn.ParallelTable()
                for _j in range(n - 1):
                    convs.add(nn.Identity())
                convs.add(resConv(nOPs[-1], nOPs[-1], 3, conv_shortcut))
                model.add(convs)


def _classifier(nIn, nLinear, avg=None):
    c = nn.Sequential()
    c.add(nn.SelectTable(1))
    if avg:
        c.add(Avg(avg, avg, 1, 1, 0, 0))
    c.add(nn.View(-1, nIn))
    c.add(nn.Linear(nIn, nLinear))
    c.add(nn.LogSoftMax())
    return c


# ------------------------------------------------------------------------------------ init (a11)
def MSRinit(model):
    """N(0, sqrt(2/(kW*kH*nOutputPlane))), bias 0 -- the builders' own fan-OUT initialiser: ConvInit of
    models/ilsvrc/rnmg.lua:288-294, MSRinit of models/cifar/nmg.lua:197-210.  (utils/modelfuncs.lua:3-11 is a different,
    fan-IN initialiser with an exponent argument: mgconv/modelfuncs.py:MSRinit.)"""
    for name in ("nn.SpatialConvolution", "cudnn.SpatialConvolution"):
        for v in model.findModules(name):
            n = v.kW * v.kH * v.nOutputPlane
            v.weight.normal_(0, math.sqrt(2.0 / n))
            v.bias.zero_()


def BNinit(model):
    """gamma = 1, beta = 0 (models/ilsvrc/rnmg.lua:295-300; utils/modelfuncs.lua:40-46)"""
    for name in ("nn.SpatialBatchNormalization", "cudnn.SpatialBatchNormalization"):
        for v in model.findModules(name):
            v.weight.fill_(1)
            v.bias.zero_()


def FCinit(model):
    """Linear bias 0 (models/ilsvrc/rnmg.lua:306-308; utils/modelfuncs.lua:34-38)"""
    for v in model.findModules("nn.Linear"):
        v.bias.zero_()


def DisableBias(model):
    """utils/modelfuncs.lua:48-54: drop conv biases (they are absorbed by the following BN)"""
    for name in ("nn.SpatialConvolution", "cudnn.SpatialConvolution"):
        for v in model.findModules(name):
            v.bias.zero_()
            v.noBias = True


# ------------------------------------------------------------------------------------ block tables
CIFAR_NMG_BLOCKS = [  # models/cifar/nmg.lua:148-154
    ([40, 40, 40], [3, 3, 3]), ([80, 40, 40], [3, 3, 3]), ([160, 80, 40], [3, 3, 3]),
    ([320, 160, 80], [3, 3, 1]), ([320, 240], [3, 1])]
CIFAR_RNMG_BLOCKS = [  # models/cifar/rnmg.lua:303-309
    ([40, 20, 10], [3, 3, 3]), ([80, 40, 20], [3, 3, 3]), ([160, 80, 40], [3, 3, 3]),
    ([320, 160, 80], [3, 3, 1]), ([320, 240], [3, 1])]
CIFAR_WIDE_BLOCKS = [  # models/cifar/prnmg.lua:330-336 (the schedule README.md:85-92 reports)
    ([64, 32, 16], [3, 3, 3]), ([128, 64, 32], [3, 3, 3]), ([256, 128, 64], [3, 3, 3]),
    ([512, 256, 128], [3, 3, 1]), ([512, 384], [3, 1])]
ILSVRC_CFG = {18: [2, 2, 2, 2], 34: [3, 4, 6, 3]}  # models/ilsvrc/rnmg.lua:244-247
ILSVRC_BLOCKS = [  # models/ilsvrc/rnmg.lua:249-255
    ([64, 32, 16], [3, 3, 3], False), ([128, 64, 32], [3, 3, 3], True),
    ([256, 128], [3, 3], True), ([512], [3], False)]


class Opt(dict):
    """stand-in for the global OPT table of opts.lua (attribute access like Lua fields)"""
    __getattr__ = dict.get
    __setattr__ = dict.__setitem__


def _nclass(opt, default=100):
    return 10 if opt.dataset == "cifar10" else default


def _finish(model, opt, net):
    if (opt.nGPU or 1) > 1:
        return makeDataParallel(model, opt.nGPU, net, bnSync=bool(opt.bnSync))
    return model


# ------------------------------------------------------------------------------------ NET tables
class BASICNET:
    """models/basic_model.lua:1-91: every hook errors until a model file overrides it; ftrain =
    forward -> criterion forward/backward -> backward (56-62); btrain = optim.sgd (64-66)."""
    name = "basic_model"

    @classmethod
    def packages(cls):
        pass

    @classmethod
    def createModel(cls, opt):
        raise NotImplementedError("BASICNET.createModel is called; you should implement your model")

    @classmethod
    def createCriterion(cls):
        raise NotImplementedError("BASICNET.createCriterion is called; you should implement your criterion")

    @classmethod
    def ftrain(cls, inputs, labels, model, criterion):
        outputs = model.forward(inputs)
        err = criterion.forward(outputs, labels)
        gradOutputs = criterion.backward(outputs, labels)
        model.backward(inputs, gradOutputs)
        return outputs, err

    @classmethod
    def btrain(cls, parameters, feval, optimState):
        from .optim import sgd
        return sgd(feval, parameters, optimState)

    @classmethod
    def ftest(cls, inputs, labels, model, criterion):
        outputs = model.forward(inputs)
        err = criterion.forward(outputs, labels)
        return outputs, err

    feval = ftest

    @classmethod
    def gradProcessing(cls, model, modelPa, modelGradPa, currentEpoch):
        pass

    @classmethod
    def arguments(cls, cmd):
        pass

    @classmethod
    def trainRule(cls, currentEpoch, opt):
        raise NotImplementedError("BASICNET.trainRule is called; you should implement your trainRule")

    @classmethod
    def createCriterion_nll(cls):
        return nn.MultiCriterion().add(nn.ClassNLLCriterion())


class cifar_nmg(BASICNET):
    """models/cifar/nmg.lua: NMG-(5*nLayer+1)"""
    name = "cifar/nmg"
    blocks = CIFAR_NMG_BLOCKS

    @classmethod
    def createModel(cls, opt):
        model = nn.Sequential()
        nIPs = [3, 3, 3]
        nLayer = opt.nLayer or 1
        for indBlock, (nOPs, kernels) in enumerate(opt.blocks or cls.blocks, 1):
            for indLayer in range(1, nLayer + 1):
                if indBlock == 1 and indLayer == 1:
                    model.add(mgConvInput_pyramid(nOPs, 3, 1e-3))
                else:
                    model.add(mgConv_plain(nIPs, nOPs, kernels, 1e-3))
                nIPs = list(nOPs)
                if indLayer == nLayer:
                    model.add(mgPool(nIPs, kernels[-1] == 1))
        model.add(_classifier(nIPs[0], _nclass(opt)))
        MSRinit(model)  # BN gamma keeps the torch7 default U(0,1): only convs are re-initialised (197-210)
        return model

    createCriterion = BASICNET.createCriterion_nll

    @classmethod
    def trainRule(cls, currentEpoch, opt):  # nmg.lua:257-263
        delta, start = 3, 1
        return {"LR": 10 ** -((currentEpoch - 1) * delta / (opt.nEpochs - 1) + start), "WD": 5e-4}


class cifar_rnmg(BASICNET):
    """models/cifar/rnmg.lua: R-NMG-(10*nLayer+2)"""
    name = "cifar/rnmg"
    blocks = CIFAR_RNMG_BLOCKS

    @classmethod
    def createModel(cls, opt):
        model = nn.Sequential()
        nIPs = [3, 3, 3]
        nLayer = opt.nLayer or 2
        for indBlock, (nOPs, kernels) in enumerate(opt.blocks or cls.blocks, 1):
            for indLayer in range(1, nLayer + 1):
                if indBlock == 1 and indLayer == 1:
                    model.add(mgConvInput_cifar_rnmg(nOPs))
                else:
                    model.add(mgConv(nIPs, nOPs, kernels))
                nIPs = list(nOPs)
                if indLayer == nLayer:
                    model.add(mgPool(nIPs, kernels[-1] == 1))
        model.add(_classifier(nIPs[0], _nclass(opt)))
        MSRinit(model); BNinit(model); FCinit(model)  # rnmg.lua:351-371
        return _finish(model, opt, cls)

    createCriterion = BASICNET.createCriterion_nll

    @classmethod
    def trainRule(cls, currentEpoch, opt):  # rnmg.lua:431-451: 0.1, x0.2 at 60/120/160
        lr = 0.1 * 0.2 ** sum(currentEpoch > e for e in (60, 120, 160))
        return {"LR": lr, "WD": 5e-4}


class cifar_prnmg(BASICNET):
    """models/cifar/prnmg.lua: progressive residual multigrid (PR-NMG)"""
    name = "cifar/prnmg"
    blocks = CIFAR_WIDE_BLOCKS

    @classmethod
    def createModel(cls, opt):
        model = nn.Sequential()
        nIPs = [3, 3, 3]
        nLayer = opt.nLayer or 2
        for indBlock, (nOPs, kernels) in enumerate(opt.blocks or cls.blocks, 1):
            if indBlock == 1:
                MultiGridsInput(model, nOPs, nLayer, 3)
                nIPs = list(nOPs)
            else:
                for _ in range(nLayer):  # MultiGrids, prnmg.lua:309-315
                    model.add(mgConv(nIPs, nOPs, kernels))
                    nIPs = list(nOPs)
            model.add(mgPool(nIPs, kernels[-1] == 1))
        model.add(_classifier(nIPs[0], _nclass(opt)))
        MSRinit(model); BNinit(model); FCinit(model)
        return _finish(model, opt, cls)

    createCriterion = BASICNET.createCriterion_nll
    trainRule = cifar_rnmg.trainRule


def mgConv_pnmg(nInputPlanes, nOutputPlanes, kernels, dropout=None):
    """models/cifar/pnmg.lua:84-112: Sequential{ ResampleConcat, ParallelTable{ [Dropout] Conv BN(1e-3) ReLU } }"""
    assert len(nInputPlanes) == len(nOutputPlanes), "number of input grid should be equal to output grid"
    assert len(nInputPlanes) == len(kernels), "should provide kernel size for every scale of grid"
    mg_conv = nn.Sequential()
    resample_concat, _nIPs = ResampleConcat(nInputPlanes)
    mg_conv.add(resample_concat)
    convs = nn.ParallelTable()
    for i in range(len(_nIPs)):
        _conv = nn.Sequential()
        if dropout:
            _conv.add(nn.Dropout(dropout))   # pnmg.lua:25-27: dropout BEFORE the convolution
        ConvBNReLU(_conv, _nIPs[i], nOutputPlanes[i], kernels[i], 1e-3)
        convs.add(_conv)
    mg_conv.add(convs)
    return mg_conv


class cifar_pnmg(BASICNET):
    """models/cifar/pnmg.lua: plain progressive multigrid (P-NMG)"""
    name = "cifar/pnmg"
    blocks = CIFAR_WIDE_BLOCKS          # pnmg.lua:244-250
    dropouts = [None, 0.1, 0.2, 0.3, 0.4]

    @classmethod
    def createModel(cls, opt):
        model = nn.Sequential()
        nIPs = [3, 3, 3]
        nLayer = opt.nLayer or 1
        for indBlock, (nOPs, kernels) in enumerate(opt.blocks or cls.blocks, 1):
            dropout = cls.dropouts[indBlock - 1] if opt.isDropout else None
            if indBlock == 1:   # MultiGridsInput, pnmg.lua:177-228
                model.add(mgConvInput_pyramid(nOPs, 3, 1e-3))
                n = len(nOPs)
                for nGrid in range(1, n + 1):
                    for _ in range(nLayer):
                        if nGrid > 1:
                            mg_convs = nn.ConcatTable()
                            for j in range(1, n - nGrid + 1):
                                mg_convs.add(nn.SelectTable(j))
                            _select = nn.ConcatTable()
                            _nOPs = []
                            for j in range(n - nGrid + 1, n + 1):
                                _select.add(nn.SelectTable(j))
                                _nOPs.append(nOPs[j - 1])
                            mg_convs.add(nn.Sequential().add(_select).add(mgConv_pnmg(_nOPs, _nOPs, [3] * len(_nOPs), dropout)))
                            model.add(mg_convs)
                            model.add(nn.FlattenTable())
                        else:
                            convs = nn.ParallelTable()
                            for _j in range(n - 1):
                                convs.add(nn.Identity())
                            _conv = nn.Sequential()
                            if dropout:
                                _conv.add(nn.Dropout(dropout))
                            convs.add(ConvBNReLU(_conv, nOPs[-1], nOPs[-1], 3, 1e-3))
                            model.add(convs)
                nIPs = list(nOPs)
            else:               # MultiGrids, pnmg.lua:230-236
                for _ in range(nLayer):
                    model.add(mgConv_pnmg(nIPs, nOPs, kernels, dropout))
                    nIPs = list(nOPs)
            model.add(mgPool(nIPs, kernels[-1] == 1))
        model.add(_classifier(nIPs[0], _nclass(opt)))
        MSRinit(model)
        return _finish(model, opt, cls)

    createCriterion = BASICNET.createCriterion_nll
    trainRule = cifar_rnmg.trainRule


class ilsvrc_rnmg(BASICNET):
    """models/ilsvrc/rnmg.lua: R-MG-18/34 -- the north-star network"""
    name = "ilsvrc/rnmg"

    @classmethod
    def createModel(cls, opt):
        inputBlock = list(opt.inputBlock or [64, 32, 16])  # (224,112,56)->(56,28,14)
        model = nn.Sequential()
        model.add(mgConvInput_ilsvrc(inputBlock))
        blocks = opt.blocks or ILSVRC_BLOCKS
        cfg = opt.cfg or ILSVRC_CFG[opt.depth or 34]
        nIPs = inputBlock
        for indBlock, (nOPs, kernels, isConcat) in enumerate(blocks, 1):
            for _ in range(cfg[indBlock - 1]):
                model.add(mgConv(nIPs, nOPs, kernels))
                nIPs = list(nOPs)
            if indBlock < len(blocks):
                model.add(mgPool(nIPs, isConcat))
        model.add(_classifier(nIPs[0], opt.nClass or 1000, avg=opt.avg if opt.avg is not None else 7))
        MSRinit(model); BNinit(model); FCinit(model)  # rnmg.lua:288-308
        return _finish(model, opt, cls)

    createCriterion = BASICNET.createCriterion_nll

    @classmethod
    def trainRule(cls, currentEpoch, opt):  # rnmg.lua:376-382
        return {"LR": 0.1 * 0.1 ** math.floor((currentEpoch - 1) / 30), "WD": 1e-4}


class mnist_prnmg(BASICNET):
    """models/mnist-cluttered/prnmg.mnist.lua: dense-prediction PR-NMG with a shrinking pyramid"""
    name = "mnist-cluttered/prnmg.mnist"

    @classmethod
    def createModel(cls, opt):
        nClass = opt.nClass or (10 if opt.dataset == "mnist-seg" else 1)  # prnmg.mnist.lua:286
        nLayer = opt.nLayer or 1
        w = opt.widths or [64, 32, 16, 8]
        blocks = [(list(w), False)] * 4 + [(w[:3], True), (w[:2], True), ([nClass], True)]  # 287-295
        model = nn.Sequential()
        nIPs = [1] * len(w)
        for indBlock, (nOPs, isDrop) in enumerate(blocks, 1):
            if indBlock == 1:
                MultiGridsInput(model, nOPs, nLayer, 1, conv_shortcut=True)
            else:
                last = indBlock == len(blocks)
                for i in range(1, nLayer + 1):  # MultiGrids 254-261 / MultiGridsOutput 263-272
                    _kernel = 1 if (last and i == nLayer) else 3
                    model.add(mgConv(nIPs, nOPs, [_kernel] * len(nOPs), isDrop if i == 1 else False,
                                     last and i == nLayer, conv_shortcut=True))
                    nIPs = list(nOPs)
            nIPs = list(nOPs)
        model.add(nn.SelectTable(1))
        model.add(nn.Sigmoid())
        MSRinit(model); BNinit(model)
        return _finish(model, opt, cls)

    @classmethod
    def createCriterion(cls):  # prnmg.mnist.lua:353-357
        return nn.MultiCriterion().add(nn.BCECriterion())

    @classmethod
    def trainRule(cls, currentEpoch, opt):
        return {"LR": 0.1 * 0.1 ** math.floor((currentEpoch - 1) / 30), "WD": 1e-4}


def mgConv_pnmg_mnist(nInputPlanes, nOutputPlanes, isDrop=False, isOut=False):
    """models/mnist-cluttered/pnmg.mnist.lua:83-121: Sequential{ ResampleConcat(nIPs, isDrop), ParallelTable{ 3x3 Conv, BN(1e-3)
    [, ReLU] } } -- mgConv (83-102) with the ReLU, mgConvOutput (104-121) without"""
    mg_conv = nn.Sequential()
    resample_concat, _nIPs = ResampleConcat(nInputPlanes, isDrop)
    mg_conv.add(resample_concat)
    convs = nn.ParallelTable()
    for i in range(len(_nIPs)):
        convs.add((ConvBN if isOut else ConvBNReLU)(nn.Sequential(), _nIPs[i], nOutputPlanes[i], 3, 1e-3))
    mg_conv.add(convs)
    return mg_conv


class mnist_pnmg(BASICNET):
    """models/mnist-cluttered/pnmg.mnist.lua: the plain progressive multigrid network for dense prediction on cluttered MNIST
    (shrinking pyramid through isDrop, last stage without ReLU, Sigmoid + BCE)"""
    name = "mnist-cluttered/pnmg.mnist"

    @classmethod
    def createModel(cls, opt):
        nClass = opt.nClass or (10 if opt.dataset == "mnist-seg" else 1)          # pnmg.mnist.lua:225
        nLayer = opt.nLayer or 1
        w = opt.widths or [64, 32, 16, 8]
        blocks = [(list(w), False)] * 4 + [(w[:3], True), (w[:2], True), ([nClass], True)]   # 226-234
        model = nn.Sequential()
        nIPs = [1] * len(w)
        for indBlock, (nOPs, isDrop) in enumerate(blocks, 1):
            if indBlock == 1:       # MultiGridsInput, 150-196
                model.add(mgConvInput_pyramid(nOPs, 1, 1e-3))
                n = len(nOPs)
                for nGrid in range(1, n + 1):
                    for _ in range(nLayer):
                        if nGrid > 1:
                            mg_convs = nn.ConcatTable()
                            for j in range(1, n - nGrid + 1):
                                mg_convs.add(nn.SelectTable(j))
                            _select = nn.ConcatTable()
                            _nOPs = []
                            for j in range(n - nGrid + 1, n + 1):
                                _select.add(nn.SelectTable(j))
                                _nOPs.append(nOPs[j - 1])
                            mg_convs.add(nn.Sequential().add(_select).add(mgConv_pnmg_mnist(_nOPs, _nOPs)))
                            model.add(mg_convs)
                            model.add(nn.FlattenTable())
                        else:
                            convs = nn.ParallelTable()
                            for _j in range(n - 1):
                                convs.add(nn.Identity())
                            convs.add(ConvBNReLU(nn.Sequential(), nOPs[-1], nOPs[-1], 3, 1e-3))
                            model.add(convs)
            else:                   # MultiGrids 198-205 / MultiGridsOutput 207-215
                last = indBlock == len(blocks)
                for i in range(1, nLayer + 1):
                    model.add(mgConv_pnmg_mnist(nIPs, nOPs, isDrop if i == 1 else False, last and i == nLayer))
                    nIPs = list(nOPs)
            nIPs = list(nOPs)
        model.add(nn.SelectTable(1))
        model.add(nn.Sigmoid())
        MSRinit(model)              # convolutions only (253-266): BN gamma keeps the Torch7 default U(0,1)
        return _finish(model, opt, cls)

    @classmethod
    def createCriterion(cls):       # pnmg.mnist.lua:275-279
        return nn.MultiCriterion().add(nn.BCECriterion())

    @classmethod
    def trainRule(cls, currentEpoch, opt):   # 312-319: LR decays log-linearly from 1e-1 to 1e-4 over nEpochs
        total = max(2, int(opt.nEpochs or 30))
        return {"LR": 10 ** -((currentEpoch - 1) * 3 / (total - 1) + 1), "WD": 5e-4}


class mnist_unmg(BASICNET):
    """models/mnist-cluttered/unmg.lua: U-Net whose every stage is a multigrid convolution; skip connections
    are zipped by nn.ConcatUnet (layers/ConcatUnet.lua) and joined per grid by MapTable(JoinTable(2))"""
    name = "mnist-cluttered/unmg"
    blocks = [([64, 32, 16], False), ([128, 64, 32], True), ([256, 128], True), ([512], None)]   # unmg.lua:181-186

    @staticmethod
    def mgUpConv(nInputPlanes, nOutputPlanes):   # unmg.lua:35-52
        assert len(nInputPlanes) == len(nOutputPlanes), "number of input grid should be equal to output grid"
        upconvs = nn.ParallelTable()
        for i in range(len(nInputPlanes)):
            upconvs.add(nn.Sequential()
                        .add(nn.SpatialFullConvolution(nInputPlanes[i], nOutputPlanes[i], 2, 2, 2, 2, 0, 0))
                        .add(SBatchNorm(nOutputPlanes[i], 1e-3)).add(ReLU(True)))
        return upconvs

    @staticmethod
    def mgConv(nIn, nOut, isReLU):               # unmg.lua:54-109: 3x3 ConvBNReLU, or 1x1 ConvBN for the output maps
        k = 3 if isReLU else 1
        return mgConv_plain(nIn, nOut, [k] * len(nIn), 1e-3, isReLU)

    @staticmethod
    def mgPool(nIn, isDrop):                     # unmg.lua:131-148: drops the coarsest grid when isDrop (mutates nIn)
        mg_pool = nn.ConcatTable()
        n = len(nIn)
        for i in range(1, n + 1):
            if i == n and isDrop:
                del nIn[i - 1]
            else:
                mg_pool.add(nn.Sequential().add(nn.SelectTable(i)).add(Max(2, 2, 2, 2, 0, 0).ceil()))
        return mg_pool

    @classmethod
    def createModel(cls, opt):
        blocks = opt.blocks or cls.blocks
        nClass = opt.nClass or (10 if (opt.dataset or "mnist-seg") == "mnist-seg" else 1)
        state = {"nIP": [1]}

        def Unet(depth):                         # unmg.lua:189-234
            unetIP = state["nIP"]
            unetOP, isDrop = blocks[depth - 1]
            model = nn.Sequential()
            if depth == len(blocks):
                model.add(cls.mgConv(unetIP, unetOP, True))
                model.add(cls.mgUpConv(unetOP, unetIP))
            else:
                if depth > 1:
                    model.add(cls.mgConv(unetIP, unetOP, True))
                else:
                    model.add(mgConvInput_pyramid(unetOP, 1, 1e-3))
                state["nIP"] = list(unetOP)
                shortcut_subnet = nn.ConcatTable()
                mg_pool = cls.mgPool(state["nIP"], isDrop)
                subnet, subnetOP = Unet(depth + 1)
                shortcut_subnet.add(nn.Identity())
                shortcut_subnet.add(nn.Sequential().add(mg_pool).add(subnet))
                model.add(shortcut_subnet)
                model.add(nn.ConcatUnet())
                model.add(nn.MapTable(nn.JoinTable(2)))
                sumOP = [(unetOP[i] if i < len(unetOP) else 0) + (subnetOP[i] if i < len(subnetOP) else 0)
                         for i in range(max(len(unetOP), len(subnetOP)))]
                model.add(cls.mgConv(sumOP, unetOP, True))
                if depth > 1:
                    model.add(cls.mgUpConv(unetOP, unetIP))
                else:
                    model.add(cls.mgConv(unetOP, [nClass] * len(unetOP), False))
                    model.add(nn.SelectTable(1))
            return model, unetIP

        model, _ = Unet(1)
        model.add(nn.Sigmoid())
        MSRinit(model)   # SpatialConvolution only (unmg.lua:239-252): up-convolutions and BN keep their defaults
        return model

    @classmethod
    def createCriterion(cls):
        return nn.MultiCriterion().add(nn.BCECriterion())

    @classmethod
    def trainRule(cls, currentEpoch, opt):
        return {"LR": 0.1 * 0.1 ** math.floor((currentEpoch - 1) / 30), "WD": 1e-4}


NETS = {c.name: c for c in (cifar_nmg, cifar_rnmg, cifar_pnmg, cifar_prnmg, ilsvrc_rnmg, mnist_prnmg, mnist_pnmg, mnist_unmg)}


def load_net(netType):
    """model.lua:21-24: NETOBJ = models/<netType>.lua with BASICNETOBJ as __index"""
    if netType not in NETS:
        raise KeyError(f"unknown -netType {netType!r}; lowered builders: {sorted(NETS)}")
    return NETS[netType]
