"""B200 tests of the input / output steps either side of the generator: bit-exact against
the reference formulas (getDepthMask as restated in synthetic.depth_masks; tensor2img as in codes/utils/util.py)."""
import numpy as np
import pytest
import torch

from depth_aware_endoscopy_sr_b200.synthetic import depth_masks as ref_depth_masks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(3, 16, 24), (2, 64, 64), (1, 135, 240)])
def test_depth_masks_are_bit_exact(shape):
    from depth_aware_endoscopy_sr_b200 import io as bio
    B, h, w = shape
    g = torch.Generator().manual_seed(B * 1000 + h)
    depth = 0.01 + 9.99 * torch.rand(B, 1, h, w, generator=g)
    depth[0, 0, 0, :4] = depth[0].max()                   # several pixels at the (exclusive) upper edge
    ref = ref_depth_masks(depth, 10)                       # the reference's torch fp32 expression, per image
    masks, labels = bio.depth_masks(depth.cuda(), 10)
    assert torch.equal(masks.cpu(), ref)
    lab_ref = torch.where(ref.sum(1) > 0, ref.argmax(1), torch.full_like(ref.argmax(1), 255)).to(torch.uint8)
    assert torch.equal(labels.cpu(), lab_ref)
    assert (labels.cpu() == 255).sum().item() >= 1        # the maximum belongs to no bin
    # depthFixedRange=True: bins over [0,1] with python-double edges
    d01 = torch.rand(B, 1, h, w, generator=g)
    m2, _ = bio.depth_masks(d01.cuda(), 10, fixed_range=True)
    exp = torch.stack([((d01[:, 0] >= (0 + 0.1 * i)) & (d01[:, 0] < (0 + 0.1 * (i + 1)))).float() for i in range(10)], 1)
    assert torch.equal(m2.cpu(), exp)


def test_tensor2img_matches_reference_conversion():
    from depth_aware_endoscopy_sr_b200 import io as bio
    g = torch.Generator().manual_seed(0)
    sr = torch.rand(2, 3, 40, 56, generator=g) * 1.4 - 0.2         # values outside [0,1] are clamped
    sr.view(-1)[:512] = (torch.arange(512) // 2).float() / 255.0 + (torch.arange(512) % 2) * 0.5 / 255.0   # exact .5 ties
    out = bio.tensor2img(sr.cuda()).cpu().numpy()
    for b in range(2):
        t = sr[b].clone().clamp_(0, 1)
        t = (t - 0) / (1 - 0)
        ref = np.transpose(t.numpy()[[2, 1, 0], :, :], (1, 2, 0))
        ref = (ref * 255.0).round().astype(np.uint8)
        assert np.array_equal(out[b], ref)
    one = bio.tensor2img(sr[0].cuda())
    assert one.shape == (40, 56, 3) and np.array_equal(one.cpu().numpy(), out[0])


def test_psnr_and_ssim_match_the_reference_formulas():
    import math
    import torch.nn.functional as F
    from depth_aware_endoscopy_sr_b200 import io as bio
    g = torch.Generator().manual_seed(3)
    gt = torch.rand(2, 3, 72, 100, generator=g)
    sr = (gt + 0.05 * torch.randn(gt.shape, generator=g)).clamp(0, 1)
    # PSNR as train.py:251-257 computes it from the tensor2img frames with a border crop of `scale` pixels
    sr8, gt8 = bio.tensor2img(sr.cuda()), bio.tensor2img(gt.cuda())
    got = bio.psnr(sr8, gt8, crop=8).cpu().numpy()
    for f in range(2):
        a = sr8[f].cpu().numpy().astype(np.float64)[8:-8, 8:-8, :]
        b = gt8[f].cpu().numpy().astype(np.float64)[8:-8, 8:-8, :]
        ref = 20 * math.log10(255.0 / math.sqrt(np.mean((a - b) ** 2)))
        assert abs(got[f] - ref) <= 1e-9 * ref
    assert math.isinf(bio.psnr(gt8, gt8).cpu()[0].item())
    # SSIM: pytorch_ssim._ssim restated with torch ops (window = outer product of the normalised 1-D Gaussian)
    w1 = torch.tensor([math.exp(-(x - 5) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
    w1 = (w1 / w1.sum()).unsqueeze(1)
    win = w1.mm(w1.t()).float().unsqueeze(0).unsqueeze(0).expand(3, 1, 11, 11).contiguous().double()
    x, y = sr.double(), gt.double()
    mu1, mu2 = F.conv2d(x, win, padding=5, groups=3), F.conv2d(y, win, padding=5, groups=3)
    s1 = F.conv2d(x * x, win, padding=5, groups=3) - mu1 ** 2
    s2 = F.conv2d(y * y, win, padding=5, groups=3) - mu2 ** 2
    s12 = F.conv2d(x * y, win, padding=5, groups=3) - mu1 * mu2
    smap = ((2 * mu1 * mu2 + 0.01 ** 2) * (2 * s12 + 0.03 ** 2)) / ((mu1 ** 2 + mu2 ** 2 + 0.01 ** 2) * (s1 + s2 + 0.03 ** 2))
    got = bio.ssim(sr.cuda(), gt.cuda(), size_average=False).cpu().double()
    ref = smap.mean(dim=(1, 2, 3))
    assert (got - ref).abs().max().item() <= 2e-5
    assert abs(bio.ssim(sr.cuda(), gt.cuda()).item() - smap.mean().item()) <= 2e-5
