"""B200-native DepthNet hot path of CUHK-AIM-Group/Depth-Aware-Endoscopy-SR.

Public surface (mirrors the reference's own, codes/models/networks.py:15-59):

    define_G(opt)      -> DepthNet         same opt dict / yml keys as the reference
    DepthNet(...)      nn.Module with the reference's constructor signature, state_dict layout and
                       forward(input, depthMap, depthMask)
    install(reference_networks_module)     patch `models.networks.define_G` so that codes/train.py and
                                           codes/test.py run unmodified on the B200 kernels

All arithmetic runs in libdasr_b200.so (hand-written sm_100a CUDA, C ABI in include/dasr.h).
"""
from .arch import DepthNet, SEAN, Encoder, Depth_Residual_Block_Mask, Classic_Residual_Block  # noqa: F401
from .networks import define_G, install  # noqa: F401

__all__ = ["DepthNet", "SEAN", "Encoder", "Depth_Residual_Block_Mask", "Classic_Residual_Block", "define_G",
           "install"]
