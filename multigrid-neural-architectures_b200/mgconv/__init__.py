"""mgconv: B200-native multigrid convolution behind the reference's Torch7 model-builder API.

    from mgconv import builders, nn          # mirrors models/*.lua and the nn vocabulary they use
    net = builders.load_net("ilsvrc/rnmg")   # -netType
    model = net.createModel(builders.Opt(depth=34, nGPU=1)).cuda()

Importing the package loads libmgconv.so (ffi.py) and fails loudly when it is missing.
"""
from . import ffi  # noqa: F401
