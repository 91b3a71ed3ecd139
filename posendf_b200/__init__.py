"""posendf_b200 -- B200-native (sm_100a) PoseNDF distance-field / projection engine.

Public surface (mirrors /root/reference/model/posendf.py):
    PoseNDF(opt)            nn.Module with the reference's constructor / forward / state_dict
    gradient(inputs, outs)  the reference's autograd helper
    Engine                  thin wrapper of one libpndf handle (C ABI in include/pndf.h)
    FusedAdam               the trainer's Adam(lr, weight_decay) as one kernel on the module's flat buffers
"""
from .module import PoseNDF, StructureEncoder, DFNet, BoneMLP, gradient  # noqa: F401


def __getattr__(name):
    if name == "Engine":
        from .engine import Engine
        return Engine
    if name == "FusedAdam":
        from .optim import FusedAdam
        return FusedAdam
    raise AttributeError(name)
