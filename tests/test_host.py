"""CPU tests of the host side: state_dict layout, define_G surface, the C ABI (symbols only), error behaviour."""
import ctypes
import os
import re
import warnings

import pytest
import torch

from common import ROOT, oracle


def _net(**kw):
    import depth_aware_endoscopy_sr_b200 as dasr
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return dasr.DepthNet(**kw)


@pytest.mark.parametrize("scale,which,latent", [(8, range(14), 256), (4, range(14), 256), (2, range(16), 32)])
def test_state_dict_layout_is_the_references(scale, which, latent):
    """Names, shapes and order of state_dict() equal the reference's (layout restated in oracle.state_layout and
    pinned to the real reference through tests/golden/*.npz sd_checksum)."""
    net = _net(which_ResBlk_depth=list(which), scale=scale, depth_latent_ch=latent, nb=16, nf=64, depthRangeNum=10)
    layout = oracle.state_layout(scale=scale, nb=16, which=tuple(which), latent=latent, K=10)
    sd = net.state_dict()
    assert sorted(sd.keys()) == sorted(layout.keys())
    for k, shape in layout.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    if scale == 8:
        assert len(sd) == 498 and sum(v.numel() for v in sd.values()) == 14795971
    assert "depth-residual14.conv1.0.weight" in sd or scale == 2


def test_define_g_reads_the_reference_opt_dict():
    import depth_aware_endoscopy_sr_b200 as dasr
    opt = {"network_G": {"which_model_G": "DepthNet", "in_nc": 3, "out_nc": 3, "nf": 64, "nb": 16, "upscale": 8,
                         "code_length": 10, "depth_latent_ch": 256, "norm_type": "weight_norm",
                         "use_trainable_params": True, "norm_gamma": 0, "norm_beta": 0,
                         "which_ResBlk_depth": list(range(14)), "ablate_depth_matrix": False,
                         "ablate_depth_block": False},
           "datasets": {"train": {"depthMaskNum": 10}}}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = dasr.define_G(opt)
    assert isinstance(net, dasr.DepthNet) and net.scale == 8 and len(net.state_dict()) == 498
    opt["datasets"] = {"test_1": {"depthMaskNum": 10}}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert dasr.define_G(opt).depthRangeNum == 10
    opt["network_G"]["which_model_G"] = "RRDBNet"
    with pytest.raises(NotImplementedError):
        dasr.define_G(opt)


def test_unsupported_configurations_fail_loudly():
    with pytest.raises(NotImplementedError):
        _net(which_ResBlk_depth=[0], scale=8, norm_type="instance_norm")
    with pytest.raises(NotImplementedError):
        _net(which_ResBlk_depth=[0], scale=8, ablate_depth_block=True)


def test_library_exports_every_declared_symbol():
    """The C-ABI shared library loads and exports every function include/dasr.h declares (no compute calls)."""
    from depth_aware_endoscopy_sr_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    header = open(os.path.join(ROOT, "include", "dasr.h")).read()
    declared = set(re.findall(r"\b(dasr_[a-z0-9_]+)\s*\(", header))
    declared -= {"dasr_status"}
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert set(_lib.EXPORTED) <= declared | {"dasr_last_error"}
    assert lib.dasr_version() >= 100


def test_cpu_tensors_raise_instead_of_falling_back():
    net = _net(which_ResBlk_depth=list(range(14)), scale=8)
    x = torch.zeros(1, 3, 16, 16)
    with torch.no_grad(), pytest.raises(RuntimeError):
        net(x, torch.zeros(1, 1, 16, 16), torch.zeros(1, 10, 16, 16))


def test_synthetic_masks_follow_get_depth_mask():
    """depth_masks restates getDepthMask (reference data/LQGTker_Depth_dataset.py:204-226): one-hot, exclusive,
    upper edge open (the arg-max pixel may fall in no bin)."""
    from depth_aware_endoscopy_sr_b200.synthetic import depth_masks
    g = torch.Generator().manual_seed(0)
    d = 0.01 + 9.99 * torch.rand(2, 1, 8, 8, generator=g)
    m = depth_masks(d, 10)
    assert m.shape == (2, 10, 8, 8) and set(m.unique().tolist()) <= {0.0, 1.0}
    assert (m.sum(1) <= 1).all() and (m.sum(1) == 1).float().mean() > 0.95


def test_multi_device_data_parallel_is_refused_with_the_way_out():
    """nn.DataParallel over several devices would share one Engine between the replicas: the replicate
    hook raises and names the one-process-per-GPU path instead."""
    import warnings
    import pytest
    import depth_aware_endoscopy_sr_b200 as dasr
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = dasr.DepthNet(which_ResBlk_depth=[0], scale=8, nb=4)
    with pytest.raises(RuntimeError, match="one process per GPU"):
        net._replicate_for_data_parallel()


def test_block_scale_of_the_blocks_behind_the_upsamplers():
    """Engine.block_scale: the feature-map resolution (relative to the LR input) of every block forward() runs
    (sftmd_arch.py:923-944): the trunk at x1, block nb-2 behind upscale1 (only x8 has one), block nb-1 behind upscale2
    (x8 and x4) -- what decides whether a depth-guided block resizes depth map and masks (normalization.py:58-59)."""
    import warnings
    import depth_aware_endoscopy_sr_b200 as dasr
    expect = {8: (2, 4), 4: (1, 2), 2: (1, 1), 3: (1, 1)}
    for scale, (s14, s15) in expect.items():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            net = dasr.DepthNet(which_ResBlk_depth=list(range(16)), scale=scale, nb=16, depth_latent_ch=32)
        eng = net.engine()
        assert [eng.block_scale(i) for i in range(13)] == [1] * 13
        assert (eng.block_scale(14), eng.block_scale(15)) == (s14, s15)
        # the blocks above LR resolution are the 32-channel ones
        assert net.block(14).nf == (32 if scale == 8 else 64)
        assert net.block(15).nf == (32 if scale >= 4 else 64)
