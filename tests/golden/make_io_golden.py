"""Golden vectors for the steps either side of the generator, produced by executing the
REAL reference functions in the build container (/root/reference does not exist on the GPU box):

    getDepthMask      codes/data/LQGTker_Depth_dataset.py:204-226   (method of LQGTKerDepthDataset, called unbound)
    tensor2img        codes/utils/util.py:566-590
    calculate_psnr    codes/utils/util.py:646-653, with the border crop of codes/train.py:251-257
    ssim              codes/pytorch_ssim/__init__.py:65-72

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_io_golden.py      ->  tests/golden/io_golden.npz

Inputs are regenerated from seeds by the tests (tests/common.py: io_case_inputs); only outputs are stored."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference/codes")
sys.dont_write_bytecode = True
sys.modules.setdefault("lmdb", types.ModuleType("lmdb"))          # not installed; only the LMDB reader needs it

from data.LQGTker_Depth_dataset import LQGTKerDepthDataset  # noqa: E402  (the reference)
from utils import util as ref_util  # noqa: E402
import pytorch_ssim as ref_ssim  # noqa: E402

from common import IO_DEPTH_SHAPES, io_depth_input, io_frames_input  # noqa: E402


def main():
    out = {}
    for (B, h, w) in IO_DEPTH_SHAPES:
        depth, d01 = io_depth_input(B, h, w)
        for tag, src, fixed in (("range", depth, False), ("fixed", d01, True)):
            masks = torch.stack([LQGTKerDepthDataset.getDepthMask(None, src[b], depthFixedRange=fixed, depthMaskNum=10)
                                 for b in range(B)], 0)          # [B,10,h,w] fp32 one-hot, as the dataset builds it
            assert masks.shape == (B, 10, h, w)
            lab = torch.where(masks.sum(1) > 0, masks.argmax(1), torch.full((B, h, w), 255)).to(torch.uint8)
            assert float(masks.sum(1).max()) <= 1.0
            out["labels_%s_%dx%dx%d" % (tag, B, h, w)] = lab.numpy()
    sr, gt = io_frames_input()
    imgs = np.stack([ref_util.tensor2img(sr[b]) for b in range(sr.shape[0])], 0)      # [B,H,W,3] uint8 BGR
    gts = np.stack([ref_util.tensor2img(gt[b]) for b in range(gt.shape[0])], 0)
    out["tensor2img_sr"] = imgs
    crop = 8
    out["psnr_crop8"] = np.array([ref_util.calculate_psnr(imgs[b][crop:-crop, crop:-crop, :], gts[b][crop:-crop, crop:-crop, :])
                                  for b in range(imgs.shape[0])])
    out["psnr_identical_is_inf"] = np.array([ref_util.calculate_psnr(gts[0], gts[0])])
    x, y = sr.clamp(0, 1), gt
    out["ssim_per_frame"] = ref_ssim.ssim(x, y, size_average=False).numpy()
    out["ssim_mean"] = np.array([ref_ssim.ssim(x, y).item()])
    path = os.path.join(HERE, "io_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KB", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
