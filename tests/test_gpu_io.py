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
