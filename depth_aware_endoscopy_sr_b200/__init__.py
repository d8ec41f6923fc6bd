"""B200-native DepthNet hot path of CUHK-AIM-Group/Depth-Aware-Endoscopy-SR.

Public surface (mirrors the reference's own, codes/models/networks.py:15-59):

    define_G(opt)      -> DepthNet         same opt dict / yml keys as the reference
    DepthNet(...)      nn.Module with the reference's constructor signature, state_dict layout and
                       forward(input, depthMap, depthMask)
    install(reference_networks_module)     patch `models.networks.define_G` so that codes/train.py and
                                           codes/test.py run unmodified on the B200 kernels
    L1Loss, dynamic_weight_mask_loss       the reference's training criteria (loss.py)
    FusedAdam                              torch.optim.Adam over one flat buffer (optim.py)
    FlatDataParallel, install_ddp          the data-parallel wrapper the reference applies when opt['dist'] (parallel.py)
    TrainStep                              optimize_parameters() as one device-resident step (trainer.py)
    depth_masks, tensor2img, psnr, ssim    getDepthMask / tensor2img / validation metrics on the device (io.py)

All arithmetic runs in libdasr_b200.so (hand-written sm_100a CUDA, C ABI in include/dasr.h).
"""
from .arch import DepthNet, SEAN, Encoder, Depth_Residual_Block_Mask, Classic_Residual_Block  # noqa: F401
from .networks import define_G, install  # noqa: F401
from .loss import L1Loss, dynamic_weight_mask_loss, training_loss  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .parallel import FlatDataParallel, install_ddp, shard_frames  # noqa: F401
from .trainer import TrainStep  # noqa: F401
from .io import depth_masks, tensor2img, psnr, ssim  # noqa: F401

__all__ = ["DepthNet", "SEAN", "Encoder", "Depth_Residual_Block_Mask", "Classic_Residual_Block", "define_G",
           "install", "L1Loss", "dynamic_weight_mask_loss", "training_loss", "FusedAdam", "FlatDataParallel",
           "install_ddp", "shard_frames", "TrainStep", "depth_masks", "tensor2img", "psnr", "ssim"]
