"""Shared helpers of the test-suite (case table, golden loading, oracle driver)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from depth_aware_endoscopy_sr_b200.synthetic import fill_state_dict, synthetic_inputs  # noqa: E402
from oracle import depthnet_oracle as oracle  # noqa: E402  (the checker; tests may import it)

GOLDEN = os.path.join(ROOT, "tests", "golden")
INIT_CASES = ["x8_b1_64_init", "x8_b2_32_init"]   # the reference's own random init (strict tolerance)
CASES = ["x8_b2_16", "x8_b1_64", "x8_b1_24x40", "x4_b1_24", "x2_b1_32", "x3_b1_24"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    scale, latent, B, h, w, seed, stride = [int(v) for v in z["meta"]]
    which = tuple(int(v) for v in z["which"])
    init = str(z["init"]) if "init" in z.files else "synthetic"
    return z, dict(scale=scale, latent=latent, B=B, h=h, w=w, seed=seed, stride=stride, which=which, init=init)


def case_tensors(meta, with_gt=True):
    if meta.get("init", "synthetic") == "default":
        # the reference's own init: our DepthNet consumes the RNG exactly like the reference's constructor
        # (checked against the reference in make_golden.py via sd_checksum)
        import warnings
        from depth_aware_endoscopy_sr_b200.arch import DepthNet
        torch.manual_seed(meta["seed"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            net = DepthNet(which_ResBlk_depth=list(meta["which"]), scale=meta["scale"], depth_latent_ch=meta["latent"],
                           nb=16, nf=64, depthRangeNum=10)
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    else:
        layout = oracle.state_layout(scale=meta["scale"], nb=16, which=meta["which"], latent=meta["latent"], K=10)
        sd = fill_state_dict(layout, seed=meta["seed"])
    inp = synthetic_inputs(meta["B"], meta["h"], meta["w"], scale=meta["scale"], seed=meta["seed"], with_gt=with_gt)
    return sd, inp


def psnr(a, b):
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return 99.0 if mse == 0 else 10.0 * np.log10(1.0 / mse)
