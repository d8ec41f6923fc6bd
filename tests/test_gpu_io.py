"""B200 tests of the steps either side of the generator against golden vectors recorded
by executing the REAL reference functions (tests/golden/make_io_golden.py -> tests/golden/io_golden.npz):
``getDepthMask`` (codes/data/LQGTker_Depth_dataset.py:204-226), ``tensor2img`` / ``calculate_psnr``
(codes/utils/util.py:566-590,646-653) and ``pytorch_ssim.ssim`` (codes/pytorch_ssim/__init__.py:65-72).
Integer / byte outputs are bit-exact; PSNR to 1e-9 relative (float64), SSIM to 2e-5 (fp32 window sums)."""
import math
import os

import numpy as np
import pytest
import torch

from common import GOLDEN, IO_DEPTH_SHAPES, io_depth_input, io_frames_input

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "io_golden.npz"))


@pytest.mark.parametrize("shape", IO_DEPTH_SHAPES)
def test_depth_masks_are_bit_exact_against_the_reference_function(shape, gold):
    from depth_aware_endoscopy_sr_b200 import io as bio
    B, h, w = shape
    depth, d01 = io_depth_input(B, h, w)
    for tag, src, fixed in (("range", depth, False), ("fixed", d01, True)):
        ref_lab = torch.from_numpy(gold["labels_%s_%dx%dx%d" % (tag, B, h, w)])
        masks, labels = bio.depth_masks(src.cuda(), 10, fixed_range=fixed)
        assert torch.equal(labels.cpu(), ref_lab), tag
        ref_masks = torch.stack([(ref_lab == k).float() for k in range(10)], 1)     # one-hot planes of the label map
        assert torch.equal(masks.cpu(), ref_masks), tag
    assert (torch.from_numpy(gold["labels_range_%dx%dx%d" % (B, h, w)]) == 255).sum().item() >= 1   # the maximum is in no bin


def test_tensor2img_is_bit_exact_against_the_reference_function(gold):
    from depth_aware_endoscopy_sr_b200 import io as bio
    sr, _gt = io_frames_input()
    out = bio.tensor2img(sr.cuda()).cpu().numpy()
    assert out.dtype == np.uint8 and np.array_equal(out, gold["tensor2img_sr"])
    one = bio.tensor2img(sr[0].cuda())
    assert one.shape == (72, 100, 3) and np.array_equal(one.cpu().numpy(), gold["tensor2img_sr"][0])


def test_psnr_and_ssim_match_the_reference_functions(gold):
    from depth_aware_endoscopy_sr_b200 import io as bio
    sr, gt = io_frames_input()
    sr8, gt8 = bio.tensor2img(sr.cuda()), bio.tensor2img(gt.cuda())
    got = bio.psnr(sr8, gt8, crop=8).cpu().numpy()          # the border crop of train.py:251-257
    np.testing.assert_allclose(got, gold["psnr_crop8"], rtol=1e-9)
    assert math.isinf(gold["psnr_identical_is_inf"][0]) and math.isinf(bio.psnr(gt8, gt8).cpu()[0].item())
    x = sr.clamp(0, 1)
    per = bio.ssim(x.cuda(), gt.cuda(), size_average=False).cpu().numpy()
    assert np.abs(per - gold["ssim_per_frame"]).max() <= 2e-5
    assert abs(bio.ssim(x.cuda(), gt.cuda()).item() - gold["ssim_mean"][0]) <= 2e-5


@pytest.mark.parametrize("B", [3, 1])       # 3: conv_output on CTA pairs, 1: the single-CTA kernel
def test_infer_frames_builds_masks_and_uint8_frames_on_the_device(B):
    """net.infer_frames(LQ, Depth): getDepthMask runs on the device in front of the generator and tensor2img inside the
    store of its output convolution (dasr_conv_out9_frames) -- the frames equal tensor2img(net(LQ, Depth, host-built
    masks)) bit for bit, only 4 of the 14 input planes are uploaded and the fp32 frames are never written."""
    import warnings
    import depth_aware_endoscopy_sr_b200 as dasr
    from depth_aware_endoscopy_sr_b200 import io as bio
    from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
    torch.manual_seed(2)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = dasr.DepthNet(which_ResBlk_depth=list(range(14)), scale=8, nb=16).cuda().eval()
    lq, depth, masks = [t.cuda() for t in synthetic_inputs(B, 24, 40, scale=8, seed=6)]
    with torch.no_grad():
        ref = bio.tensor2img(net(lq, depth, masks))
    frames = net.infer_frames(lq, depth)
    assert frames.dtype == torch.uint8 and tuple(frames.shape) == (B, 192, 320, 3)
    assert torch.equal(frames, ref)
    sr = net.infer_frames(lq, depth, out="float")
    with torch.no_grad():
        assert torch.equal(sr, net(lq, depth, masks))


@pytest.mark.parametrize("shape,ratio", [((2, 10, 16, 16), (2, 2)), ((1, 1, 24, 40), (4, 4)), ((3, 10, 9, 7), (2, 4))])
def test_nearest_resize_equals_torch_interpolate(shape, ratio):
    """dasr_nearest_up == F.interpolate(mode='nearest') -- what a SEAN instance above LR resolution applies to the depth
    map and the masks (codes/models/modules/normalization.py:58-59); bit-exact (a copy)."""
    from depth_aware_endoscopy_sr_b200 import _lib as L
    B, C, h, w = shape
    g = torch.Generator().manual_seed(h * 100 + w)
    x = torch.randn(B, C, h, w, generator=g).cuda()
    H, W = h * ratio[0], w * ratio[1]
    out = torch.empty(B, C, H, W, device="cuda")
    L.check(L.load().dasr_nearest_up(L.ptr(x), L.ptr(out), B * C, h, w, H, W, L.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(out, torch.nn.functional.interpolate(x, size=(H, W), mode="nearest"))
