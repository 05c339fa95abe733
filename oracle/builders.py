"""CPU ORACLE (test infrastructure, NOT product code) -- restatement of the
reference's Lua model builders on PyTorch-CPU through oracle/t7nn.py.

PARITY UNPINNED: Torch7 cannot run in this environment and the reference ships no
tests/golden vectors; the only numeric anchors it publishes are parameter / MAC
counts (README.md:85-92,109) and stage-shape comments (models/ilsvrc/rnmg.lua:241,
251-254), all checked in tests/test_oracle.py.

Each function cites the reference lines it follows.  Only tests/, smoke() and
bench.py's cpu_baseline / --impl reference leg may import this module.
"""
import torch
from . import t7nn as nn

Convolution = nn.SpatialConvolution
ReLU = nn.ReLU
Max = nn.SpatialMaxPooling
Avg = nn.SpatialAveragePooling
SBatchNorm = nn.SpatialBatchNormalization
UpSample = nn.SpatialUpSamplingNearest
UpConvolution = nn.SpatialFullConvolution


# ------------------------------------------------------------------ basic units
def Shortcut(nIP, nOP, allow_conv=False):
    """models/ilsvrc/rnmg.lua:13-20; conv variant models/mnist-cluttered/prnmg.mnist.lua:13-25"""
    if nOP > nIP:
        return nn.Padding(1, nOP - nIP, 3)
    elif nIP > nOP:
        assert allow_conv
        conv = nn.Sequential()
        conv.add(Convolution(nIP, nOP, 1, 1, 1, 1, 0, 0))
        conv.add(SBatchNorm(nOP))
        return conv
    return nn.Identity()


def ConvBNReLU(mod, nIP, nOP, kernel, eps=1e-5):
    """models/ilsvrc/rnmg.lua:22-30 (eps default); models/cifar/nmg.lua:18-29 (eps 1e-3)"""
    k = kernel
    p = 0 if k == 1 else 1
    mod.add(Convolution(nIP, nOP, k, k, 1, 1, p, p))
    mod.add(SBatchNorm(nOP, eps))
    mod.add(ReLU(True))
    return mod


def ConvBN(mod, nIP, nOP, kernel, eps=1e-5):
    """models/ilsvrc/rnmg.lua:32-39"""
    k = kernel
    p = 0 if k == 1 else 1
    mod.add(Convolution(nIP, nOP, k, k, 1, 1, p, p))
    mod.add(SBatchNorm(nOP, eps))
    return mod


def ResampleConcat(nIPs, isDrop=False):
    """models/ilsvrc/rnmg.lua:41-89; isDrop: models/mnist-cluttered/prnmg.mnist.lua:44-92"""
    resample_concat = nn.ConcatTable()
    nOPs = []
    nGrids = len(nIPs) - 1 if isDrop else len(nIPs)
    for iG in range(1, nGrids + 1):
        grid = nn.Sequential()
        multi_scales = nn.ConcatTable()
        nIP = 0
        if iG - 1 > 0:
            finer_scale = nn.Sequential()
            finer_scale.add(nn.SelectTable(iG - 1))
            finer_scale.add(Max(2, 2, 2, 2, 0, 0).ceil())
            multi_scales.add(finer_scale)
            nIP += nIPs[iG - 2]
        multi_scales.add(nn.SelectTable(iG))
        nIP += nIPs[iG - 1]
        if iG + 1 <= nGrids:
            coarser_scale = nn.Sequential()
            coarser_scale.add(nn.SelectTable(iG + 1))
            coarser_scale.add(UpSample(2))
            multi_scales.add(coarser_scale)
            nIP += nIPs[iG]
        grid.add(multi_scales)
        grid.add(nn.JoinTable(2))
        resample_concat.add(grid)
        nOPs.append(nIP)
    return resample_concat, nOPs


def plain_mgConv(nInputPlanes, nOutputPlanes, kernels, eps=1e-3, relu=None):
    """models/cifar/nmg.lua:31-86 (relu always); models/mnist-cluttered/unmg.lua:54-109
    (isReLU false => 1x1 ConvBN)"""
    assert len(nInputPlanes) == len(nOutputPlanes) == len(kernels)
    rc, nIPs = ResampleConcat(nInputPlanes)
    multi_grids = nn.ConcatTable()
    for iG in range(len(nInputPlanes)):
        grid = nn.Sequential()
        grid.add(rc.mods[iG])
        if relu is None or relu:
            ConvBNReLU(grid, nIPs[iG], nOutputPlanes[iG], kernels[iG], eps)
        else:
            ConvBN(grid, nIPs[iG], nOutputPlanes[iG], kernels[iG], eps)
        multi_grids.add(grid)
    return multi_grids


def res_mgConv(nInputPlanes, nOutputPlanes, kernels, isDrop=False, isOut=False, conv_shortcut=False):
    """models/ilsvrc/rnmg.lua:91-159 (= cifar/rnmg.lua:102-173, cifar/prnmg.lua:122-193);
    isDrop/isOut: models/mnist-cluttered/prnmg.mnist.lua:108-175"""
    if not isDrop:
        assert len(nInputPlanes) == len(nOutputPlanes)
    mg_conv = nn.Sequential()
    shortcut_convs = nn.ConcatTable()
    convs = nn.Sequential()
    resample_concat, _nIPs = ResampleConcat(nInputPlanes, isDrop)
    convs.add(resample_concat)
    conv_bn_relu = nn.ParallelTable()
    for i in range(len(_nIPs)):
        conv_bn_relu.add(ConvBNReLU(nn.Sequential(), _nIPs[i], nOutputPlanes[i], kernels[i]))
    convs.add(conv_bn_relu)
    resample_concat, _nIPs = ResampleConcat(nOutputPlanes, False)
    convs.add(resample_concat)
    conv_bn = nn.ParallelTable()
    for i in range(len(_nIPs)):
        conv_bn.add(ConvBN(nn.Sequential(), _nIPs[i], nOutputPlanes[i], kernels[i]))
    convs.add(conv_bn)
    shortcut_convs.add(convs)

    nShortcut = len(_nIPs)
    shortcut = nn.ConcatTable()
    for i in range(nShortcut):
        _sc = nn.Sequential()
        _sc.add(nn.SelectTable(i + 1))
        _sc.add(Shortcut(nInputPlanes[i], nOutputPlanes[i], conv_shortcut))
        shortcut.add(_sc)
    shortcut_convs.add(shortcut)

    add_shortcut_convs = nn.ConcatTable()
    for i in range(nShortcut):
        pick = nn.ConcatTable()
        pick.add(nn.Sequential().add(nn.SelectTable(1)).add(nn.SelectTable(i + 1)))
        pick.add(nn.Sequential().add(nn.SelectTable(2)).add(nn.SelectTable(i + 1)))
        s = nn.Sequential()
        s.add(pick).add(nn.CAddTable(True))
        if not isOut:
            s.add(ReLU(True))
        add_shortcut_convs.add(s)
    mg_conv.add(shortcut_convs)
    mg_conv.add(add_shortcut_convs)
    return mg_conv


def resConv(nIP, nOP, kernel, conv_shortcut=False):
    """models/cifar/prnmg.lua:102-120; models/mnist-cluttered/prnmg.mnist.lua:94-106"""
    s = nn.Sequential()
    ConvBNReLU(s, nIP, nOP, kernel)
    ConvBN(s, nOP, nOP, kernel)
    return nn.Sequential() \
        .add(nn.ConcatTable().add(s).add(Shortcut(nIP, nOP, conv_shortcut))) \
        .add(nn.CAddTable(True)) \
        .add(ReLU(True))


def mgPool(nInputPlanes, isConcat):
    """models/ilsvrc/rnmg.lua:191-224 (same in every cifar file). MUTATES nInputPlanes."""
    mg_pool = nn.ConcatTable()
    nGrids = len(nInputPlanes)
    for i in range(1, nGrids + 1):
        proc = nn.Sequential()
        if i == nGrids - 1 and isConcat:
            pool_cat = nn.ConcatTable()
            pool = nn.Sequential().add(nn.SelectTable(i)).add(Max(2, 2, 2, 2, 0, 0).ceil())
            pool_cat.add(pool)
            pool_cat.add(nn.SelectTable(i + 1))
            proc.add(pool_cat)
            proc.add(nn.JoinTable(2))
            nInputPlanes[i - 1] = nInputPlanes[i - 1] + nInputPlanes[i]
            del nInputPlanes[i]
        else:
            proc.add(nn.SelectTable(i))
            proc.add(Max(2, 2, 2, 2, 0, 0).ceil())
        mg_pool.add(proc)
        if i == nGrids - 1 and isConcat:
            break
    return mg_pool


def image_pyramid_convs(nOutputPlanes, nIn=3, eps=1e-5):
    """CIFAR/MNIST mgConvInput: [AvgPool r] -> 3x3 conv -> BN -> ReLU per grid
    (models/cifar/nmg.lua:88-106 eps 1e-3; cifar/prnmg.lua:195-213; prnmg.mnist.lua:177-195)"""
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


def classifier(nIn, nLinear, avg=None):
    c = nn.Sequential()
    c.add(nn.SelectTable(1))
    if avg:
        c.add(Avg(avg, avg, 1, 1, 0, 0))
    c.add(nn.View(-1, nIn))
    c.add(nn.Linear(nIn, nLinear))
    c.add(nn.LogSoftMax())
    return c


# ------------------------------------------------------------------ models
CIFAR_NARROW = [  # models/cifar/nmg.lua:148-154
    ([40, 40, 40], [3, 3, 3]), ([80, 40, 40], [3, 3, 3]), ([160, 80, 40], [3, 3, 3]),
    ([320, 160, 80], [3, 3, 1]), ([320, 240], [3, 1])]
CIFAR_RNMG_NARROW = [  # models/cifar/rnmg.lua:303-309
    ([40, 20, 10], [3, 3, 3]), ([80, 40, 20], [3, 3, 3]), ([160, 80, 40], [3, 3, 3]),
    ([320, 160, 80], [3, 3, 1]), ([320, 240], [3, 1])]
CIFAR_WIDE = [  # models/cifar/prnmg.lua:330-336 (README's table matches this schedule)
    ([64, 32, 16], [3, 3, 3]), ([128, 64, 32], [3, 3, 3]), ([256, 128, 64], [3, 3, 3]),
    ([512, 256, 128], [3, 3, 1]), ([512, 384], [3, 1])]


def cifar_nmg(nLayer=1, nClass=100, blocks=None):
    """models/cifar/nmg.lua:143-213 (NMG-(5*nLayer+1)); BN gamma keeps torch7 default U(0,1)
    because only convs are re-initialised (197-210)."""
    blocks = blocks or CIFAR_NARROW
    model = nn.Sequential()
    nIPs = [3, 3, 3]
    for indBlock, (nOPs, kernels) in enumerate(blocks, 1):
        for indLayer in range(1, nLayer + 1):
            if indBlock == 1 and indLayer == 1:
                model.add(image_pyramid_convs(nOPs, 3, 1e-3))
            else:
                model.add(plain_mgConv(nIPs, nOPs, kernels, 1e-3))
            nIPs = list(nOPs)
            if indLayer == nLayer:
                model.add(mgPool(nIPs, kernels[-1] == 1))
    model.add(classifier(nIPs[0], nClass))
    nn.conv_init_msr_fanout(model)
    return model


def cifar_rnmg_input(nOutputPlanes):
    """models/cifar/rnmg.lua:175-254: image pyramid convs + one residual unit (all 3x3)"""
    m = nn.Sequential()
    m.add(image_pyramid_convs(nOutputPlanes, 3))
    unit = res_mgConv(nOutputPlanes, nOutputPlanes, [3] * len(nOutputPlanes))
    for sub in unit:
        m.add(sub)
    return m


def cifar_rnmg(nLayer=2, nClass=100, blocks=None):
    """models/cifar/rnmg.lua:298-386 (R-NMG-(10*nLayer+2))"""
    blocks = blocks or CIFAR_RNMG_NARROW
    model = nn.Sequential()
    nIPs = [3, 3, 3]
    for indBlock, (nOPs, kernels) in enumerate(blocks, 1):
        for indLayer in range(1, nLayer + 1):
            if indBlock == 1 and indLayer == 1:
                model.add(cifar_rnmg_input(nOPs))
            else:
                model.add(res_mgConv(nIPs, nOPs, kernels))
            nIPs = list(nOPs)
            if indLayer == nLayer:
                model.add(mgPool(nIPs, kernels[-1] == 1))
    model.add(classifier(nIPs[0], nClass))
    nn.conv_init_msr_fanout(model)
    nn.bn_init(model)
    nn.linear_bias_zero(model)
    return model


def _progressive_input(model, nOPs, nLayer, nIn, conv_shortcut=False):
    """MultiGridsInput: models/cifar/prnmg.lua:258-307; prnmg.mnist.lua:205-252"""
    model.add(image_pyramid_convs(nOPs, nIn))
    n = len(nOPs)
    for nGrid in range(1, n + 1):
        if nGrid > 1:
            for _ in range(nLayer):
                mg_convs = nn.ConcatTable()
                for j in range(1, n - nGrid + 1):
                    mg_convs.add(nn.SelectTable(j))
                _mg_conv = nn.Sequential()
                _select = nn.ConcatTable()
                _nOPs = []
                for j in range(n - nGrid + 1, n + 1):
                    _select.add(nn.SelectTable(j))
                    _nOPs.append(nOPs[j - 1])
                _mg_conv.add(_select)
                _mg_conv.add(res_mgConv(_nOPs, _nOPs, [3] * len(_nOPs), conv_shortcut=conv_shortcut))
                mg_convs.add(_mg_conv)
                model.add(mg_convs)
                model.add(nn.FlattenTable())
        else:
            for _ in range(nLayer):
                convs = nn.ParallelTable()
                for j in range(n - 1):
                    convs.add(nn.Identity())
                convs.add(resConv(nOPs[-1], nOPs[-1], 3, conv_shortcut))
                model.add(convs)


def cifar_prnmg(nLayer=2, nClass=100, blocks=None):
    """models/cifar/prnmg.lua:323-386 (PR-NMG)"""
    blocks = blocks or CIFAR_WIDE
    model = nn.Sequential()
    nIPs = [3, 3, 3]
    for indBlock, (nOPs, kernels) in enumerate(blocks, 1):
        if indBlock == 1:
            _progressive_input(model, nOPs, nLayer, 3)
            nIPs = list(nOPs)
        else:
            for _ in range(nLayer):  # MultiGrids, prnmg.lua:309-315
                model.add(res_mgConv(nIPs, nOPs, kernels))
                nIPs = list(nOPs)
        model.add(mgPool(nIPs, kernels[-1] == 1))
    model.add(classifier(nIPs[0], nClass))
    nn.conv_init_msr_fanout(model)
    nn.bn_init(model)
    nn.linear_bias_zero(model)
    return model


def cifar_pnmg(nLayer=1, nClass=100, blocks=None):
    """models/cifar/pnmg.lua:238-310 (plain progressive multigrid; BN eps 1e-3, gamma keeps U(0,1))"""
    blocks = blocks or CIFAR_WIDE
    model = nn.Sequential()
    nIPs = [3, 3, 3]
    for indBlock, (nOPs, kernels) in enumerate(blocks, 1):
        if indBlock == 1:  # MultiGridsInput, pnmg.lua:177-228
            model.add(image_pyramid_convs(nOPs, 3, 1e-3))
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
                        mg_convs.add(nn.Sequential().add(_select).add(plain_mgConv(_nOPs, _nOPs, [3] * len(_nOPs), 1e-3)))
                        model.add(mg_convs)
                        model.add(nn.FlattenTable())
                    else:
                        convs = nn.ParallelTable()
                        for _j in range(n - 1):
                            convs.add(nn.Identity())
                        convs.add(ConvBNReLU(nn.Sequential(), nOPs[-1], nOPs[-1], 3, 1e-3))
                        model.add(convs)
            nIPs = list(nOPs)
        else:
            for _ in range(nLayer):
                model.add(plain_mgConv(nIPs, nOPs, kernels, 1e-3))
                nIPs = list(nOPs)
        model.add(mgPool(nIPs, kernels[-1] == 1))
    model.add(classifier(nIPs[0], nClass))
    nn.conv_init_msr_fanout(model)
    return model


def ilsvrc_stem(nOutputPlanes):
    """mgConvInput, models/ilsvrc/rnmg.lua:161-189"""
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


ILSVRC_CFG = {18: [2, 2, 2, 2], 34: [3, 4, 6, 3]}  # models/ilsvrc/rnmg.lua:244-247
ILSVRC_BLOCKS = [  # models/ilsvrc/rnmg.lua:249-255
    ([64, 32, 16], [3, 3, 3], False), ([128, 64, 32], [3, 3, 3], True),
    ([256, 128], [3, 3], True), ([512], [3], False)]


def ilsvrc_rnmg(depth=34, nClass=1000, input_block=None, blocks=None, cfg=None, avg=7):
    """models/ilsvrc/rnmg.lua:236-323 (R-MG-18/34). input_block/blocks/cfg/avg overridable
    for reduced-size test cases."""
    inputBlock = list(input_block or [64, 32, 16])
    blocks = blocks or ILSVRC_BLOCKS
    cfg = cfg or ILSVRC_CFG[depth]
    model = nn.Sequential()
    model.add(ilsvrc_stem(inputBlock))
    nIPs = inputBlock
    for indBlock, (nOPs, kernels, isConcat) in enumerate(blocks, 1):
        for _ in range(cfg[indBlock - 1]):
            model.add(res_mgConv(nIPs, nOPs, kernels))
            nIPs = list(nOPs)
        if indBlock < len(blocks):
            model.add(mgPool(nIPs, isConcat))
    model.add(classifier(nIPs[0], nClass, avg=avg))
    nn.conv_init_msr_fanout(model)
    nn.bn_init(model)
    nn.linear_bias_zero(model)
    return model


def mnist_prnmg(nLayer=1, nClass=1):
    """models/mnist-cluttered/prnmg.mnist.lua:281-340 (+ MultiGrids 254-261, MultiGridsOutput 263-272)"""
    blocks = [([64, 32, 16, 8], False)] * 4 + [([64, 32, 16], True), ([64, 32], True), ([nClass], True)]
    model = nn.Sequential()
    nIPs = [1, 1, 1, 1]
    for indBlock, (nOPs, isDrop) in enumerate(blocks, 1):
        if indBlock == 1:
            _progressive_input(model, nOPs, nLayer, 1, conv_shortcut=True)
        else:
            last = indBlock == len(blocks)
            for i in range(1, nLayer + 1):
                _drop = isDrop if i == 1 else False
                _kernel = 1 if (last and i == nLayer) else 3
                _isOut = last and i == nLayer
                model.add(res_mgConv(nIPs, nOPs, [_kernel] * len(nOPs), _drop, _isOut, conv_shortcut=True))
                nIPs = list(nOPs)
        nIPs = list(nOPs)
    model.add(nn.SelectTable(1))
    model.add(nn.Sigmoid())
    nn.conv_init_msr_fanout(model)
    nn.bn_init(model)
    return model


def pnmg_mnist_mgConv(nInputPlanes, nOutputPlanes, isDrop=False, isOut=False):
    """models/mnist-cluttered/pnmg.mnist.lua:83-102 (mgConv) / 104-121 (mgConvOutput: ConvBN, no ReLU)"""
    mg_conv = nn.Sequential()
    rc, nIPs = ResampleConcat(nInputPlanes, isDrop)
    mg_conv.add(rc)
    convs = nn.ParallelTable()
    for i in range(len(nIPs)):
        convs.add((ConvBN if isOut else ConvBNReLU)(nn.Sequential(), nIPs[i], nOutputPlanes[i], 3, 1e-3))
    mg_conv.add(convs)
    return mg_conv


def mnist_pnmg(nLayer=1, nClass=1):
    """models/mnist-cluttered/pnmg.mnist.lua:219-272 (+ MultiGridsInput 150-196, MultiGrids 198-205, MultiGridsOutput 207-215)"""
    blocks = [([64, 32, 16, 8], False)] * 4 + [([64, 32, 16], True), ([64, 32], True), ([nClass], True)]
    model = nn.Sequential()
    nIPs = [1, 1, 1, 1]
    for indBlock, (nOPs, isDrop) in enumerate(blocks, 1):
        if indBlock == 1:
            model.add(image_pyramid_convs(nOPs, 1, 1e-3))
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
                        mg_convs.add(nn.Sequential().add(_select).add(pnmg_mnist_mgConv(_nOPs, _nOPs)))
                        model.add(mg_convs)
                        model.add(nn.FlattenTable())
                    else:
                        convs = nn.ParallelTable()
                        for _j in range(n - 1):
                            convs.add(nn.Identity())
                        convs.add(ConvBNReLU(nn.Sequential(), nOPs[-1], nOPs[-1], 3, 1e-3))
                        model.add(convs)
        else:
            last = indBlock == len(blocks)
            for i in range(1, nLayer + 1):
                model.add(pnmg_mnist_mgConv(nIPs, nOPs, isDrop if i == 1 else False, last and i == nLayer))
                nIPs = list(nOPs)
        nIPs = list(nOPs)
    model.add(nn.SelectTable(1))
    model.add(nn.Sigmoid())
    nn.conv_init_msr_fanout(model)
    return model


def mnist_unmg(nClass=10):
    """models/mnist-cluttered/unmg.lua:174-258 (U-MG; the nn.ConcatUnet user)"""
    blocks = [([64, 32, 16], False), ([128, 64, 32], True), ([256, 128], True), ([512], None)]

    def mgUpConv(nIn, nOut):  # unmg.lua:42-52
        up = nn.ParallelTable()
        for i in range(len(nIn)):
            mod = nn.Sequential()
            mod.add(UpConvolution(nIn[i], nOut[i], 2, 2, 2, 2, 0, 0))
            mod.add(SBatchNorm(nOut[i], 1e-3))
            mod.add(ReLU(True))
            up.add(mod)
        return up

    def umgConv(nIn, nOut, isReLU):  # unmg.lua:54-109: 3x3 ConvBNReLU or 1x1 ConvBN
        k = 3 if isReLU else 1
        return plain_mgConv(nIn, nOut, [k] * len(nIn), 1e-3, relu=isReLU)

    def umgPool(nIn, isDrop):  # unmg.lua:131-148 (mutates nIn)
        mg_pool = nn.ConcatTable()
        n = len(nIn)
        for i in range(1, n + 1):
            if i == n and isDrop:
                del nIn[i - 1]
            else:
                mg_pool.add(nn.Sequential().add(nn.SelectTable(i)).add(Max(2, 2, 2, 2, 0, 0).ceil()))
        return mg_pool

    state = {"nIP": [1]}

    def Unet(depth):  # unmg.lua:189-234
        unetIP = state["nIP"]
        unetOP, isDrop = blocks[depth - 1]
        model = nn.Sequential()
        if depth == len(blocks):
            model.add(umgConv(unetIP, unetOP, True))
            model.add(mgUpConv(unetOP, unetIP))
        else:
            if depth > 1:
                model.add(umgConv(unetIP, unetOP, True))
            else:
                model.add(image_pyramid_convs(unetOP, 1, 1e-3))
            state["nIP"] = list(unetOP)
            shortcut_subnet = nn.ConcatTable()
            mg_pool = umgPool(state["nIP"], isDrop)
            subnet, subnetOP = Unet(depth + 1)
            shortcut_subnet.add(nn.Identity())
            shortcut_subnet.add(nn.Sequential().add(mg_pool).add(subnet))
            model.add(shortcut_subnet)
            model.add(nn.ConcatUnet())
            model.add(nn.MapTable(nn.JoinTable(2)))
            sumOP = [(unetOP[i] if i < len(unetOP) else 0) + (subnetOP[i] if i < len(subnetOP) else 0)
                     for i in range(max(len(unetOP), len(subnetOP)))]
            model.add(umgConv(sumOP, unetOP, True))
            if depth > 1:
                model.add(mgUpConv(unetOP, unetIP))
            else:
                model.add(umgConv(unetOP, [nClass] * 3, False))
                model.add(nn.SelectTable(1))
        return model, unetIP

    model, _ = Unet(1)
    model.add(nn.Sigmoid())
    nn.conv_init_msr_fanout(model)  # MSRinit on SpatialConvolution only (unmg.lua:239-252)
    return model


# ------------------------------------------------------------------ structural known answers
def count_params(model):
    return sum(p.numel() for p in model.parameters())


def count_conv_macs(model, x):
    """MACs = sum over convs of Cin*Cout*k*k*Hout*Wout for one image"""
    macs = [0]
    hooks = []

    def hook(m, inp, out):
        k = m.kernel_size[0] * m.kernel_size[1]
        if isinstance(m, torch.nn.ConvTranspose2d):
            macs[0] += m.in_channels * m.out_channels * k * inp[0].shape[2] * inp[0].shape[3]
        else:
            macs[0] += m.in_channels * m.out_channels * k * out.shape[2] * out.shape[3]
    for m in model.modules():
        if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
            hooks.append(m.register_forward_hook(hook))
    model.eval()
    with torch.no_grad():
        model(x)
    for h in hooks:
        h.remove()
    model.train()
    return macs[0]
