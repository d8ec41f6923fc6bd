"""DepthNet / SEAN / depth-guided blocks: the reference's generator surface on top of libdasr_b200.so.

What this file mirrors (paths relative to /root/reference/codes):

* ``DepthNet``                    models/modules/sftmd_arch.py:837-950
* ``Encoder``                     models/modules/sftmd_arch.py:735-783   (+ RegionWiseAvgPooling 709-733)
* ``Depth_Residual_Block_Mask``   models/modules/sftmd_arch.py:808-834
* ``Classic_Residual_Block``      models/modules/sftmd_arch.py:128-151
* ``SEAN``                        models/modules/normalization.py:7-92

The classes below are *parameter containers*: they register exactly the reference's children and parameters
(same names, shapes, registration order and RNG consumption, so ``state_dict()`` has the reference's 498 keys
and ``torch.manual_seed(s); DepthNet(...)`` initialises to the same values).  None of their ``nn.Conv2d`` /
``weight_norm`` children is ever *called*: ``DepthNet.forward`` hands the raw parameters to the CUDA kernels
behind the C ABI of ``include/dasr.h`` (``engine.py``).  There is no PyTorch / CPU fallback -- CPU tensors raise.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from . import engine as _engine


def _wn(m: nn.Module) -> nn.Module:
    # legacy weight_norm on purpose: it owns the `weight_g` / `weight_v` keys of the reference checkpoints
    # (torch.nn.utils.parametrizations.weight_norm would rename them).  sftmd_arch.py:740,851
    return torch.nn.utils.weight_norm(m)


def _conv3(cin: int, cout: int, stride: int = 1) -> nn.Conv2d:
    return nn.Conv2d(cin, cout, 3, stride=stride, padding=1)


class RegionWiseAvgPooling(nn.Module):
    """Parameter-free; kept so that ``encoder.pool`` exists like in the reference (sftmd_arch.py:709-733)."""

    def __init__(self):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)


class Encoder(nn.Module):
    """sftmd_arch.py:735-783, ``norm_type == 'weight_norm'`` branch."""

    def __init__(self, in_nc: int = 3, latent_ch: int = 256, isBaseline: bool = False):
        super().__init__()
        self.isBaseline = isBaseline
        self.actvn = nn.LeakyReLU(0.2, False)
        self.layer1 = _wn(_conv3(in_nc, 32))
        self.layer2 = _wn(_conv3(32, 64, stride=2))
        self.layer3 = _wn(_conv3(64, 128, stride=2))
        self.layer4 = _wn(nn.ConvTranspose2d(128, latent_ch, 3, stride=2, padding=1))
        self.layer5 = _wn(_conv3(latent_ch, latent_ch, stride=2))
        self.pool = RegionWiseAvgPooling()


class SEAN(nn.Module):
    """Depth-mask spatially-adaptive normalisation, default (inject_st, no ablation) configuration of
    normalization.py:7-49."""

    def __init__(self, label_nc: int = 10, norm_nc: int = 32, len_latent: int = 256, use_trainable_params: bool = True,
                 norm_gamma: float = 0.1, norm_beta: float = 0.1):
        super().__init__()
        self.len_latent = len_latent
        self.label_nc = label_nc
        self.norm_nc = norm_nc
        self.param_free_norm = nn.InstanceNorm2d(norm_nc, affine=False)
        self.A_i_j = nn.Conv2d(label_nc, label_nc, kernel_size=1, padding=0)
        self.mlp_gamma_s = nn.Conv2d(len_latent, norm_nc, kernel_size=3, padding=1)
        self.mlp_beta_s = nn.Conv2d(len_latent, norm_nc, kernel_size=3, padding=1)
        if use_trainable_params:
            self.alpha_beta = Parameter(torch.rand(1), requires_grad=True)
            self.alpha_gamma = Parameter(torch.rand(1), requires_grad=True)
        else:  # fixed blend factors from the yml (normalization.py:33-35); kept out of the state_dict
            self.register_buffer("alpha_beta", torch.tensor([float(norm_beta)]), persistent=False)
            self.register_buffer("alpha_gamma", torch.tensor([float(norm_gamma)]), persistent=False)
        self.mlp_mask = nn.Sequential(nn.Conv2d(1, 2 * norm_nc, kernel_size=3, padding=1), nn.ReLU())
        self.mlp_gamma_o = nn.Conv2d(2 * norm_nc, norm_nc, kernel_size=3, padding=1)
        self.mlp_beta_o = nn.Conv2d(2 * norm_nc, norm_nc, kernel_size=3, padding=1)


class Depth_Residual_Block_Mask(nn.Module):
    """Depth-guided block (DGB), sftmd_arch.py:808-834."""

    def __init__(self, nf: int = 64, depth_latent_ch: int = 256, depthRangeNum: int = 10,
                 use_trainable_params: bool = True, norm_gamma: float = 0.1, norm_beta: float = 0.1):
        super().__init__()
        self.nf = nf
        kw = dict(label_nc=depthRangeNum, norm_nc=nf, len_latent=depth_latent_ch,
                  use_trainable_params=use_trainable_params, norm_gamma=norm_gamma, norm_beta=norm_beta)
        # construction order (RNG) and registration order differ in the reference; both are reproduced
        conv1 = [_conv3(nf, nf), nn.InstanceNorm2d(nf, affine=False)]
        self.norm1 = SEAN(**kw)
        self.actv1 = nn.ReLU(True)
        conv2 = [_conv3(nf, nf), nn.InstanceNorm2d(nf, affine=False)]
        self.norm2 = SEAN(**kw)
        self.conv1 = nn.Sequential(*conv1)
        self.conv2 = nn.Sequential(*conv2)


class Classic_Residual_Block(nn.Module):
    """sftmd_arch.py:128-151, weight_norm branch."""

    def __init__(self, nf: int = 64):
        super().__init__()
        self.nf = nf
        self.block = nn.Sequential(_wn(_conv3(nf, nf)), nn.ReLU(True), _wn(_conv3(nf, nf)))


class DepthNet(nn.Module):
    """Drop-in for ``models.modules.sftmd_arch.DepthNet`` (constructor signature of sftmd_arch.py:838,
    ``forward(input, depthMap, depthMask)`` of sftmd_arch.py:912-950).

    Inputs are fp32 NCHW CUDA tensors like the reference's; the output is fp32 ``[B,3,s*h,s*w]`` in [0,1].
    Internally activations are NHWC bf16 and every op is a kernel of libdasr_b200.so.
    """

    def __init__(self, which_ResBlk_depth=(), in_nc=3, out_nc=3, nf=64, nb=16, scale=4, input_para=10, min=0.0,
                 max=1.0, depth_latent_ch=256, depthRangeNum=10, norm_type="weight_norm", use_trainable_params=True,
                 norm_gamma=0.1, norm_beta=0.1, ablate_depth_matrix=False, ablate_depth_block=False):
        super().__init__()
        if norm_type != "weight_norm":
            raise NotImplementedError("only norm_type='weight_norm' (every shipped yml) is on the B200 path")
        if ablate_depth_matrix or ablate_depth_block:
            raise NotImplementedError("ablation variants (ablate_depth_matrix / ablate_depth_block) are out of scope")
        if in_nc != 3 or out_nc != 3 or nf != 64:
            raise NotImplementedError("the B200 kernels are specialised for in_nc=out_nc=3, nf=64 (every shipped yml)")
        if scale not in (2, 3, 4, 8):
            raise NotImplementedError("scale %r: the reference builds x2, x3, x4 and x8" % (scale,))
        if use_trainable_params is None:
            use_trainable_params = True
        self.scale = scale
        self.min = min
        self.max = max
        self.para = input_para
        self.num_blocks = nb
        self.which_ResBlk_depth = list(which_ResBlk_depth or [])
        self.isBaseline = len(self.which_ResBlk_depth) == 0
        self.depth_latent_ch = depth_latent_ch
        self.depthRangeNum = depthRangeNum

        self.encoder = Encoder(in_nc=in_nc, latent_ch=depth_latent_ch, isBaseline=self.isBaseline)
        self.head = nn.Sequential(_wn(_conv3(32, 64)), nn.LeakyReLU(0.2), _wn(_conv3(64, 64)), nn.LeakyReLU(0.2))

        num_last_block = 1 if scale == 3 else int(math.log(scale, 2))
        ch_last2_upscale = 64 if scale == 4 else 32
        ch_last_upscale = 64 if scale < 4 else 32
        for i in range(nb):
            ch = 32 if i > nb - num_last_block else nf
            if i in self.which_ResBlk_depth:
                blk = Depth_Residual_Block_Mask(nf=ch, depth_latent_ch=depth_latent_ch, depthRangeNum=depthRangeNum,
                                                use_trainable_params=use_trainable_params, norm_gamma=norm_gamma,
                                                norm_beta=norm_beta)
                self.add_module("depth-residual%d" % (i + 1), blk)
            else:
                self.add_module("classic-residual%d" % (i + 1), Classic_Residual_Block(nf=ch))

        self.upscale1 = nn.Sequential(_wn(_conv3(64, 64 * 4)), nn.PixelShuffle(2), nn.LeakyReLU(0.2, inplace=True),
                                      _wn(_conv3(64, 32)), nn.LeakyReLU(0.2, inplace=True))
        self.upscale2 = nn.Sequential(_wn(_conv3(ch_last2_upscale, 32 * 4)), nn.PixelShuffle(2),
                                      nn.LeakyReLU(0.2, inplace=True), _wn(_conv3(32, 32)),
                                      nn.LeakyReLU(0.2, inplace=True))
        final_scale = 3 if scale == 3 else 2
        self.upscale3 = nn.Sequential(_wn(_conv3(ch_last_upscale, 32 * final_scale ** 2)), nn.PixelShuffle(final_scale),
                                      nn.LeakyReLU(0.2, inplace=True))
        self.conv_output = nn.Conv2d(32, out_nc, kernel_size=9, stride=1, padding=4, bias=True)
        self._engine = None

    # ------------------------------------------------------------------------------------------ blocks
    def block(self, i: int) -> nn.Module:
        """Block with 0-based index ``i`` (children are named 1-based, sftmd_arch.py:886-889)."""
        kind = "depth-residual" if i in self.which_ResBlk_depth else "classic-residual"
        return getattr(self, "%s%d" % (kind, i + 1))

    def block_order(self):
        """(index, position) of the blocks ``forward`` actually runs: the trunk loop covers 0..nb-4, then
        nb-2 after upscale1 and nb-1 after upscale2; block nb-3 is constructed but never called
        (sftmd_arch.py:923-944)."""
        nb = self.num_blocks
        return [(i, "trunk") for i in range(nb - 3)] + [(nb - 2, "up1"), (nb - 1, "up2")]

    def engine(self) -> "_engine.Engine":
        if self._engine is None:
            self._engine = _engine.Engine(self)
        return self._engine

    def _replicate_for_data_parallel(self):
        """``nn.DataParallel`` over SEVERAL devices (the reference's default wrapper when ``opt['dist']`` is false,
        F_model_depthCond.py:35) replicates the module per device and runs the replicas in threads.  A replica
        shallow-copies ``__dict__``, i.e. it would share this module's Engine -- packed weights, flat gradient
        buffers and recorded graphs bound to the parameters of device 0.  The B200 path is one process per GPU
        (``parallel.FlatDataParallel`` / ``install_ddp``: the reference's own ``--launcher pytorch`` mode), so a
        multi-device replicate is refused with the way out; a single-device DataParallel never replicates."""
        raise RuntimeError(
            "DepthNet (B200) does not run under multi-device nn.DataParallel: its engine (packed weights, flat "
            "gradient buffers, CUDA graphs) is bound to one device.  Use one process per GPU -- torchrun + "
            "codes/train.py --launcher pytorch with depth_aware_endoscopy_sr_b200.install_ddp() (INTEGRATION.md, "
            "section 3), or set gpu_ids to a single device.")

    def forward(self, input, depthMap, depthMask):
        return self.engine().forward(input, depthMap, depthMask)

    @torch.no_grad()
    def infer_frames(self, input, depthMap, depthMask=None, out: str = "uint8", depthFixedRange: bool = False):
        """LR frames + depth maps -> SR frames, with the steps the reference runs on the CPU either side of the generator
        done on the device:

        * ``depthMask=None``: the K one-hot depth masks are built from ``depthMap`` by ``io.depth_masks`` --
          ``LQGTKerDepthDataset.getDepthMask(depth, depthFixedRange, depthRangeNum)`` of the data pipeline
          (codes/data/LQGTker_Depth_dataset.py:157,204-226; bit-exact) -- so only the LR frame and its depth map
          (4 fp32 planes instead of 14) have to be uploaded;
        * ``out="uint8"``: ``util.tensor2img`` as codes/test.py:87 calls it (clamp, x255, round, uint8, RGB->BGR, HWC;
          bit-exact) runs in the store of the output convolution and the ``[B, sH, sW, 3]`` uint8 frames are returned
          -- a quarter of the bytes of the fp32 tensor, which is never written; ``out="float"`` returns what ``forward``
          returns.

        Inference only (no autograd); frames are independent, so a stream is sharded over GPUs by frame
        (``parallel.shard_frames``) with no collective."""
        from . import io as _io
        if out not in ("uint8", "float"):
            raise ValueError("out must be 'uint8' or 'float'")
        if depthMask is None:
            depthMask, _labels = _io.depth_masks(depthMap, self.depthRangeNum, fixed_range=depthFixedRange)
        # uint8: tensor2img is fused into the store of the output convolution (dasr_conv_out9_frames; bit-identical to
        # io.tensor2img of the fp32 tensor, which then never exists in HBM)
        return self.engine().infer(input, depthMap, depthMask, frames=(out == "uint8"))
