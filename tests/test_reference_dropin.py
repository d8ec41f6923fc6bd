"""Drop-in check against the REAL reference code base (CPU, no kernels run): with ``install()`` the reference's own
``create_model(opt)`` (codes/models/__init__.py:5-30 -> F_Model_depthCond.__init__, codes/models/F_model_depthCond.py
:21-129) builds around the B200 generator -- DataParallel wrap, print_network, optimizer over named_parameters,
lr scheduler -- and checkpoints move between the two implementations with ``strict=True``.

Needs /root/reference (present in the build container, absent on the GPU box): skipped otherwise."""
import logging
import os
import sys
import warnings

import pytest
import torch

REF = "/root/reference/codes"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference code base is not available here")


def _opt(scale=8):
    import yaml
    with open(os.path.join(REF, "options", "train", "train_depthNet_SEAN_depthMask_x8.yml")) as f:
        opt = yaml.safe_load(f)
    opt["is_train"] = True
    opt["gpu_ids"] = None          # -> device cpu (base_model.py:11)
    opt["dist"] = False
    opt["scale"] = scale
    opt["path"] = {"pretrain_model_G": None, "models": "/tmp", "training_state": "/tmp", "strict_load": True,
                   "resume_state": None}
    opt["network_G"]["scale"] = scale
    from options.options import dict_to_nonedict     # the reference's own "missing key -> None" dict (options.py:103)
    return dict_to_nonedict(opt)


def test_reference_create_model_builds_around_the_b200_generator():
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import models.networks as networks
            from models import create_model
            from models.modules import sftmd_arch
            import depth_aware_endoscopy_sr_b200 as dasr
            logging.getLogger("base").setLevel(logging.ERROR)
            stock = networks.define_G
            dasr.install(networks)
            try:
                opt = _opt()
                torch.manual_seed(0)
                model = create_model(opt)          # the reference's F_Model_depthCond, unmodified
            finally:
                networks.define_G = stock
            net = model.netG.module                 # DataParallel wrap of F_model_depthCond.py:35
            assert isinstance(net, dasr.DepthNet)
            # the optimiser the reference built sees every generator parameter + the 10 dynamic-loss weights
            n_opt = sum(p.numel() for g in model.optimizer_G.param_groups for p in g["params"])
            assert n_opt == sum(p.numel() for p in net.parameters()) + 10 == 14795971 + 10
            assert len(model.schedulers) == 1
            # same construction consumes the RNG identically: identical initial weights as the reference class
            gopt = opt["network_G"]
            torch.manual_seed(0)
            ref = sftmd_arch.DepthNet(gopt["which_ResBlk_depth"], in_nc=gopt["in_nc"], out_nc=gopt["out_nc"], nf=gopt["nf"],
                                      nb=gopt["nb"], scale=gopt["upscale"], input_para=gopt["code_length"],
                                      depth_latent_ch=gopt["depth_latent_ch"],
                                      depthRangeNum=opt["datasets"]["train"]["depthMaskNum"],
                                      norm_type=gopt["norm_type"], use_trainable_params=gopt["use_trainable_params"],
                                      norm_gamma=gopt["norm_gamma"], norm_beta=gopt["norm_beta"])
            sd_ref, sd_new = ref.state_dict(), net.state_dict()
            assert list(sd_ref.keys()) == list(sd_new.keys())
            for k in sd_ref:
                assert sd_ref[k].shape == sd_new[k].shape and torch.equal(sd_ref[k], sd_new[k]), k
            # checkpoints move both ways with strict=True (base_model.load_network strips a 'module.' prefix itself)
            net.load_state_dict(sd_ref, strict=True)
            ref.load_state_dict(sd_new, strict=True)
            assert [n for n, _ in ref.named_parameters()] == [n for n, _ in net.named_parameters()]
            # no CPU fallback: the reference's test() on CPU tensors must fail loudly, not compute silently
            model.var_L = torch.rand(1, 3, 8, 8)
            model.var_depth = torch.rand(1, 1, 8, 8)
            model.var_depthMask = torch.zeros(1, 10, 8, 8)
            with pytest.raises(RuntimeError):
                model.test()
    finally:
        if REF in sys.path:
            sys.path.remove(REF)
        for m in [k for k in sys.modules if k.split(".")[0] in ("models", "utils", "options")]:
            del sys.modules[m]
