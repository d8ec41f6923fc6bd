"""CPU: the oracle's restatement of getDepthMask and the synthetic-input generator against the golden vectors
recorded from the reference's own function (tests/golden/make_io_golden.py)."""
import os

import numpy as np
import pytest
import torch

from common import GOLDEN, IO_DEPTH_SHAPES, io_depth_input, oracle
from depth_aware_endoscopy_sr_b200.synthetic import depth_masks


@pytest.mark.parametrize("shape", IO_DEPTH_SHAPES)
def test_get_depth_mask_restatement_and_generator_match_reference(shape):
    gold = np.load(os.path.join(GOLDEN, "io_golden.npz"))
    B, h, w = shape
    depth, d01 = io_depth_input(B, h, w)
    for tag, src, fixed in (("range", depth, False), ("fixed", d01, True)):
        ref_lab = torch.from_numpy(gold["labels_%s_%dx%dx%d" % (tag, B, h, w)])
        ref = torch.stack([(ref_lab == k).float() for k in range(10)], 1)
        got = torch.stack([oracle.get_depth_mask(src[b], fixed, 10) for b in range(B)], 0)
        assert torch.equal(got, ref), tag
    # the vectorised generator behind synthetic_inputs (per-image range) produces the same masks
    ref_lab = torch.from_numpy(gold["labels_range_%dx%dx%d" % (B, h, w)])
    assert torch.equal(depth_masks(depth, 10), torch.stack([(ref_lab == k).float() for k in range(10)], 1))
