"""`define_G(opt)`: the drop-in boundary (reference codes/models/networks.py:15-59, DepthNet branch 41-49)."""
from __future__ import annotations

from .arch import DepthNet


def define_G(opt):
    """Same contract as the reference: reads ``opt['network_G']`` and ``depthMaskNum`` of the *first* dataset
    entry (``train`` or ``test_1``); returns the generator ``nn.Module``.  Only ``which_model_G: DepthNet`` is on
    the B200 path -- anything else raises, exactly like an unknown name does in the reference."""
    opt_net = opt["network_G"]
    which_model = opt_net["which_model_G"]
    if which_model != "DepthNet":
        raise NotImplementedError("Generator model [{:s}] not recognized".format(str(which_model)))
    datalist = list(opt["datasets"].items())
    if datalist[0][0] == "train":
        depth_range_num = opt["datasets"]["train"]["depthMaskNum"]
    else:
        depth_range_num = opt["datasets"]["test_1"]["depthMaskNum"]

    def get(key, default=None):
        try:
            v = opt_net[key]
        except KeyError:
            v = None
        return default if v is None else v

    return DepthNet(which_ResBlk_depth=get("which_ResBlk_depth", []), in_nc=opt_net["in_nc"], out_nc=opt_net["out_nc"],
                    nf=opt_net["nf"], nb=opt_net["nb"], scale=opt_net["upscale"], input_para=get("code_length", 10),
                    depth_latent_ch=get("depth_latent_ch", 256), depthRangeNum=depth_range_num,
                    norm_type=get("norm_type", "weight_norm"), use_trainable_params=get("use_trainable_params", True),
                    norm_gamma=get("norm_gamma", 0.1), norm_beta=get("norm_beta", 0.1),
                    ablate_depth_block=bool(get("ablate_depth_block", False)),
                    ablate_depth_matrix=bool(get("ablate_depth_matrix", False)))


def install(networks_module=None):
    """Route the reference's ``models.networks.define_G`` to this implementation for ``which_model_G: DepthNet``
    (other generators keep the reference's own code).  Call it before ``create_model(opt)``; see INTEGRATION.md."""
    if networks_module is None:
        import models.networks as networks_module  # the reference package, must be on sys.path
    stock = networks_module.define_G

    def routed(opt):
        if opt["network_G"]["which_model_G"] == "DepthNet":
            return define_G(opt)
        return stock(opt)

    routed.__wrapped__ = stock
    networks_module.define_G = routed
    return networks_module
