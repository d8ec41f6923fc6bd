"""CPU oracle for the DepthNet hot path -- TEST INFRASTRUCTURE ONLY.

This file restates, in plain functional PyTorch fp32/fp64 on the CPU, the algorithm of the reference
generator.  It is NOT part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product path
(``depth_aware_endoscopy_sr_b200``) never imports anything from ``oracle/`` and has no CPU fallback.

Parity pin: ``tests/golden/*.npz`` were produced by ``tests/golden/make_golden.py``, which imports the
*real* reference (``/root/reference/codes/models/modules/sftmd_arch.py``) in the build container, loads
the same seeded ``state_dict`` and records its outputs; ``tests/test_oracle_golden.py`` checks this
restatement against those vectors (fp32 round-off only).  The reference ships no golden vectors or
tests of its own, so the pin is "reference executed here", not "reference KATs".

Each function cites the reference lines it follows (paths relative to /root/reference/codes).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

EPS_IN = 1e-5  # nn.InstanceNorm2d default eps (models/modules/sftmd_arch.py:813, normalization.py:17)


# --------------------------------------------------------------------------------------------- bf16 emulation
class _RoundBF16(torch.autograd.Function):
    """Round to bf16 in the forward, identity in the backward (straight-through)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


class bf16_operands:
    """Context manager: every convolution of this module rounds its two operands (activation, weight) to bf16
    and accumulates in fp32 -- the arithmetic contract of a bf16 tensor-core implementation (what
    ``torch.autocast(bfloat16)`` does to the reference).  DepthNet's gradients are ill-conditioned (26
    normalisation layers): this 2^-9 operand perturbation alone moves early-layer gradients of the REFERENCE by
    30-40 % (tests/test_gpu_backward.py), so gradient parity of a bf16 implementation is defined against this
    oracle variant; the fp32 oracle remains the checker for outputs."""

    def __enter__(self):
        self._orig = F.conv2d
        self._orig_t = F.conv_transpose2d
        orig, orig_t = self._orig, self._orig_t

        def conv2d(x, w, b=None, stride=1, padding=0):
            return orig(_RoundBF16.apply(x), _RoundBF16.apply(w), b, stride=stride, padding=padding)

        def conv_t(x, w, b=None, stride=1, padding=0):
            return orig_t(_RoundBF16.apply(x), _RoundBF16.apply(w), b, stride=stride, padding=padding)

        F.conv2d, F.conv_transpose2d = conv2d, conv_t
        return self

    def __exit__(self, *exc):
        F.conv2d, F.conv_transpose2d = self._orig, self._orig_t
        return False


# --------------------------------------------------------------------------------------------- helpers
def weight_norm(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """torch.nn.utils.weight_norm(dim=0): w = g * v / ||v||, norm over all dims but 0
    (models/modules/sftmd_arch.py:740,851; for ConvTranspose2d dim 0 is Cin)."""
    n = v.reshape(v.shape[0], -1).norm(dim=1).reshape([-1] + [1] * (v.dim() - 1))
    return v * (g / n)


def _wn_conv(sd, prefix, x, stride=1, padding=1):
    w = weight_norm(sd[prefix + ".weight_g"], sd[prefix + ".weight_v"])
    return F.conv2d(x, w, sd[prefix + ".bias"], stride=stride, padding=padding)


def instance_norm(x: torch.Tensor) -> torch.Tensor:
    """InstanceNorm2d(affine=False): biased variance, eps inside the sqrt."""
    mu = x.mean(dim=(2, 3), keepdim=True)
    var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    return (x - mu) / torch.sqrt(var + EPS_IN)


def lrelu(x):
    return F.leaky_relu(x, 0.2)


# --------------------------------------------------------------------------------------------- activation patterns
# DepthNet is piecewise smooth: ReLU / LeakyReLU masks, the final clamp and the sign of the L1 loss switch the
# gradient discontinuously.  An implementation whose forward differs by eps flips the units within eps of a switching
# point, and each flip changes the gradient by O(1) of that unit's contribution -- so gradients of two correct
# implementations differ by ~sqrt(flipped fraction), not by eps (the reference's own fp32 gradients deviate from
# its fp64 gradients by up to 20 % on single parameters this way, tests/golden/*.npz ``grad_dev32``).  To compare
# gradients ON THE SAME SMOOTH PIECE the oracle can (a) record its activation pattern and (b) be evaluated with a
# pattern recorded elsewhere (the CUDA run): ``with activation_pattern(record=d)`` / ``activation_pattern(force=d)``.
# Keys: the prefix of the convolution in front of the activation ("head.0", "upscale2.0", "<block>.block.0"),
# "<sean>.actv", "<sean>.out" (ReLU after norm1 / after the residual add of norm2), "<classic block>.out", "clamp"
# and "l1.sign".  Without a context every function below is the plain reference arithmetic.
_PATTERN = {"record": None, "force": None}


class activation_pattern:
    def __init__(self, record: Optional[dict] = None, force: Optional[dict] = None):
        self.new = {"record": record, "force": force}

    def __enter__(self):
        self.old = dict(_PATTERN)
        _PATTERN.update(self.new)
        return self

    def __exit__(self, *exc):
        _PATTERN.update(self.old)
        return False


def _act(x, key, slope=0.0):
    """ReLU (slope 0) / LeakyReLU(slope) of the call site ``key``."""
    if _PATTERN["record"] is not None:
        _PATTERN["record"][key] = (x > 0).detach()
    m = None if _PATTERN["force"] is None else _PATTERN["force"].get(key)
    if m is None:
        return F.leaky_relu(x, slope) if slope else F.relu(x)
    m = m.to(x.dtype)
    return x * (m + slope * (1.0 - m))


def _clamp01(x, key="clamp"):
    if _PATTERN["record"] is not None:
        _PATTERN["record"][key] = ((x > 0) & (x < 1)).detach()
    m = None if _PATTERN["force"] is None else _PATTERN["force"].get(key)
    if m is None:
        return torch.clamp(x, 0.0, 1.0)
    return torch.where(m, x, torch.clamp(x, 0.0, 1.0).detach())


def _abs(d, key="l1.sign"):
    if _PATTERN["record"] is not None:
        _PATTERN["record"][key] = (d > 0).detach()
    m = None if _PATTERN["force"] is None else _PATTERN["force"].get(key)
    if m is None:
        return d.abs()
    return d * (2.0 * m.to(d.dtype) - 1.0)


# --------------------------------------------------------------------------------------------- encoder
def region_pool(feat: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """RegionWiseAvgPooling.forward (models/modules/sftmd_arch.py:714-733) -> [B,K,C]."""
    if mask.shape[2] != feat.shape[2] or mask.shape[3] != feat.shape[3]:
        mask = F.interpolate(mask, size=feat.shape[2:], mode="bilinear", align_corners=True)
        mask = (mask >= 0.5).to(feat.dtype)
    sum_feat = torch.einsum("bkhw,bchw->bkc", mask, feat)
    sum_mask = mask.sum(dim=(2, 3)).unsqueeze(2)
    return sum_feat / (sum_mask + 1e-10)


def encoder_forward(sd, x, mask, cap=None):
    """Encoder.forward (models/modules/sftmd_arch.py:771-783), weight_norm branch 743-749."""
    e1 = _wn_conv(sd, "encoder.layer1", x)
    f0 = _act(e1, "encoder.layer1", 0.2)
    e2 = _wn_conv(sd, "encoder.layer2", f0, stride=2)
    e3 = _wn_conv(sd, "encoder.layer3", _act(e2, "encoder.layer2", 0.2), stride=2)
    w4 = weight_norm(sd["encoder.layer4.weight_g"], sd["encoder.layer4.weight_v"])
    e4 = F.conv_transpose2d(_act(e3, "encoder.layer3", 0.2), w4, sd["encoder.layer4.bias"], stride=2, padding=1)
    e5 = _wn_conv(sd, "encoder.layer5", _act(e4, "encoder.layer4", 0.2), stride=2)
    vec = region_pool(e5, mask)
    if cap is not None:
        cap.update(e1=e1, e2=e2, e3=e3, e4=e4, e5=e5, depthVec=vec)
    return f0, vec


# --------------------------------------------------------------------------------------------- SEAN
def sean_gamma_beta(sd, p, depth_map, depth_mask, st, size):
    """gamma/beta of SEAN.forward (models/modules/normalization.py:58-88), literal form."""
    depth_map = F.interpolate(depth_map, size=size, mode="nearest")
    depth_mask = F.interpolate(depth_mask, size=size, mode="nearest")
    actv = _act(F.conv2d(depth_map, sd[p + ".mlp_mask.0.weight"], sd[p + ".mlp_mask.0.bias"], padding=1), p + ".actv")
    beta_o = F.conv2d(actv, sd[p + ".mlp_beta_o.weight"], sd[p + ".mlp_beta_o.bias"], padding=1)
    gamma_o = F.conv2d(actv, sd[p + ".mlp_gamma_o.weight"], sd[p + ".mlp_gamma_o.bias"], padding=1)
    # A_i_j: 1x1 conv over the label axis (normalization.py:80)
    stp = torch.einsum("ji,bic->bjc", sd[p + ".A_i_j.weight"][:, :, 0, 0], st) + sd[p + ".A_i_j.bias"][None, :, None]
    # style_map[b,c,h,w] = sum_k st'[b,k,c] mask[b,k,h,w]   (normalization.py:81-82)
    style_map = torch.einsum("bkc,bkhw->bchw", stp, depth_mask)
    beta_s = F.conv2d(style_map, sd[p + ".mlp_beta_s.weight"], sd[p + ".mlp_beta_s.bias"], padding=1)
    gamma_s = F.conv2d(style_map, sd[p + ".mlp_gamma_s.weight"], sd[p + ".mlp_gamma_s.bias"], padding=1)
    a_g = sd[p + ".alpha_gamma"]
    a_b = sd[p + ".alpha_beta"]
    gamma = a_g * gamma_s + (1.0 - a_g) * gamma_o
    beta = a_b * beta_s + (1.0 - a_b) * beta_o
    return gamma, beta, dict(actv=actv, gamma_o=gamma_o, beta_o=beta_o, gamma_s=gamma_s, beta_s=beta_s, stp=stp)


def sean_forward(sd, p, x, depth_map, depth_mask, st):
    """SEAN.forward (models/modules/normalization.py:52-92), default path (inject_st, no ablation)."""
    assert st.shape[1] == depth_mask.shape[1]
    normalized = instance_norm(x)
    gamma, beta, _ = sean_gamma_beta(sd, p, depth_map, depth_mask, st, x.shape[2:])
    return normalized * (1 + gamma) + beta


# ---- algebraic restatement of the style branch as a per-image dynamic 3x3 convolution
def style_table(w_s: torch.Tensor, stp: torch.Tensor) -> torch.Tensor:
    """T[b,o,k,t,u] = sum_c W_s[o,c,t,u] * st'[b,k,c]."""
    return torch.einsum("octu,bkc->boktu", w_s, stp)


def dynconv_apply(table: torch.Tensor, bias: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """out[b] = conv3x3(mask[b] (K ch), T[b]) + bias  -- equals mlp_*_s(style_map)."""
    outs = [F.conv2d(mask[b:b + 1], table[b], bias, padding=1) for b in range(mask.shape[0])]
    return torch.cat(outs, 0)


# --------------------------------------------------------------------------------------------- blocks
def dgb_forward(sd, p, x, depth_map, depth_mask, st, cap=None):
    """Depth_Residual_Block_Mask.forward (models/modules/sftmd_arch.py:826-834)."""
    y1 = instance_norm(F.conv2d(x, sd[p + ".conv1.0.weight"], sd[p + ".conv1.0.bias"], padding=1))
    a = _act(sean_forward(sd, p + ".norm1", y1, depth_map, depth_mask, st), p + ".norm1.out")
    y2 = instance_norm(F.conv2d(a, sd[p + ".conv2.0.weight"], sd[p + ".conv2.0.bias"], padding=1))
    z = sean_forward(sd, p + ".norm2", y2, depth_map, depth_mask, st)
    out = _act(x + z, p + ".norm2.out")
    if cap is not None:
        cap[p + ".a"] = a
        cap[p + ".out"] = out
    return out


def classic_forward(sd, p, x):
    """Classic_Residual_Block.forward, weight_norm branch (models/modules/sftmd_arch.py:131-136,147-151)."""
    f = _wn_conv(sd, p + ".block.0", x)
    f = _wn_conv(sd, p + ".block.2", _act(f, p + ".block.0"))
    return _act(x + f, p + ".out")


def _block(sd, idx, which, x, depth_map, depth_mask, vec, cap):
    if idx in which:
        return dgb_forward(sd, "depth-residual%d" % (idx + 1), x, depth_map, depth_mask, vec, cap)
    return classic_forward(sd, "classic-residual%d" % (idx + 1), x)


# --------------------------------------------------------------------------------------------- network
def depthnet_forward(sd: Dict[str, torch.Tensor], lq, depth_map, depth_mask, scale=8, nb=16,
                     which=tuple(range(14)), cap: Optional[dict] = None, clamp=True):
    """DepthNet.forward (models/modules/sftmd_arch.py:912-950)."""
    f0, vec = encoder_forward(sd, lq, depth_mask, cap)
    fea_bef = _act(_wn_conv(sd, "head.2", _act(_wn_conv(sd, "head.0", f0), "head.0", 0.2)), "head.2", 0.2)
    x = fea_bef
    for i in range(nb - 3):
        x = _block(sd, i, which, x, depth_map, depth_mask, vec, cap)
    x = x + fea_bef
    if cap is not None:
        cap["fea_bef"] = fea_bef
        cap["feat_add1"] = x
    if scale == 8:
        x = _act(F.pixel_shuffle(_wn_conv(sd, "upscale1.0", x), 2), "upscale1.0", 0.2)
        x = _act(_wn_conv(sd, "upscale1.3", x), "upscale1.3", 0.2)
    x = _block(sd, nb - 2, which, x, depth_map, depth_mask, vec, cap)
    if scale >= 4:
        x = _act(F.pixel_shuffle(_wn_conv(sd, "upscale2.0", x), 2), "upscale2.0", 0.2)
        x = _act(_wn_conv(sd, "upscale2.3", x), "upscale2.3", 0.2)
    x = _block(sd, nb - 1, which, x, depth_map, depth_mask, vec, cap)
    r = 3 if scale == 3 else 2
    x = _act(F.pixel_shuffle(_wn_conv(sd, "upscale3.0", x), r), "upscale3.0", 0.2)
    out = F.conv2d(x, sd["conv_output.weight"], sd["conv_output.bias"], padding=4)
    if cap is not None:
        cap["feat_up3"] = x
        cap["pre_clamp"] = out
    return _clamp01(out) if clamp else out


# --------------------------------------------------------------------------------------------- loss
def dynamic_mask_loss(sr, hr, masks, trainable_weight, l_w=10.0):
    """dynamic_weight_mask_loss.forward, smoothl1 branch (models/modules/mask_loss.py:64-90)."""
    sw = F.softmax(trainable_weight, dim=0)
    raw = []
    for i in range(masks.shape[1]):
        m = F.interpolate(masks[:, i:i + 1], size=sr.shape[2:], mode="nearest")
        m3 = torch.cat([m, m, m], dim=1)
        l = F.smooth_l1_loss(m3 * sr, m3 * hr, reduction="none").sum() / m3.sum()
        raw.append(l)
    weighted = sum(sw[i] * raw[i] for i in range(len(raw))) * l_w
    return raw, weighted, sw


def training_loss(sr, hr, masks, trainable_weight, l_pix_w=1.0, l_dyn_w=10.0):
    """F_Model_depthCond.optimize_parameters loss (models/F_model_depthCond.py:163-190)."""
    l_pix = l_pix_w * _abs(sr - hr).mean()
    raw, l_dyn, sw = dynamic_mask_loss(sr, hr, masks, trainable_weight, l_dyn_w)
    return l_pix + l_dyn, l_pix, l_dyn, raw


# --------------------------------------------------------------------------------------------- layout
def state_layout(scale=8, nb=16, which=tuple(range(14)), latent=256, K=10):
    """Names + shapes of DepthNet.state_dict() (models/modules/sftmd_arch.py:838-910); SURVEY 8(b)."""
    from collections import OrderedDict
    L = OrderedDict()

    def wn(p, co, ci, k=3, transposed=False):
        L[p + ".bias"] = (co,)
        if transposed:  # ConvTranspose2d weight [Cin,Cout,k,k]; weight_norm dim 0 = Cin
            L[p + ".weight_g"] = (ci, 1, 1, 1)
            L[p + ".weight_v"] = (ci, co, k, k)
        else:
            L[p + ".weight_g"] = (co, 1, 1, 1)
            L[p + ".weight_v"] = (co, ci, k, k)

    def conv(p, co, ci, k=3):
        L[p + ".weight"] = (co, ci, k, k)
        L[p + ".bias"] = (co,)

    wn("encoder.layer1", 32, 3)
    wn("encoder.layer2", 64, 32)
    wn("encoder.layer3", 128, 64)
    wn("encoder.layer4", latent, 128, transposed=True)
    wn("encoder.layer5", latent, latent)
    wn("head.0", 64, 32)
    wn("head.2", 64, 64)
    num_last = 1 if scale == 3 else int(math.log(scale, 2))
    for i in range(nb):
        ch = 32 if i > nb - num_last else 64
        if i in which:
            p = "depth-residual%d" % (i + 1)
            for j in (1, 2):
                n = "%s.norm%d" % (p, j)
                if j == 1:
                    pass
            # registration order in the reference: norm1, norm2, conv1, conv2 (sftmd_arch.py:814-824)
            for j in (1, 2):
                n = "%s.norm%d" % (p, j)
                L[n + ".alpha_beta"] = (1,)
                L[n + ".alpha_gamma"] = (1,)
                conv(n + ".A_i_j", K, K, 1)
                conv(n + ".mlp_gamma_s", ch, latent)
                conv(n + ".mlp_beta_s", ch, latent)
                conv(n + ".mlp_mask.0", 2 * ch, 1)
                conv(n + ".mlp_gamma_o", ch, 2 * ch)
                conv(n + ".mlp_beta_o", ch, 2 * ch)
            conv(p + ".conv1.0", ch, ch)
            conv(p + ".conv2.0", ch, ch)
        else:
            p = "classic-residual%d" % (i + 1)
            wn(p + ".block.0", ch, ch)
            wn(p + ".block.2", ch, ch)
    ch2 = 64 if scale == 4 else 32
    ch3 = 64 if scale < 4 else 32
    r = 3 if scale == 3 else 2
    wn("upscale1.0", 256, 64)
    wn("upscale1.3", 32, 64)
    wn("upscale2.0", 128, ch2)
    wn("upscale2.3", 32, 32)
    wn("upscale3.0", 32 * r * r, ch3)
    conv("conv_output", 3, 32, 9)
    return L


# --------------------------------------------------------------------------------------------- input preparation
def get_depth_mask(depth_map: torch.Tensor, fixed_range: bool = True, num: int = 10) -> torch.Tensor:
    """``LQGTKerDepthDataset.getDepthMask`` (data/LQGTker_Depth_dataset.py:204-226) for ONE depth map [1,h,w] or
    [h,w]: ``num`` one-hot fp32 planes [num,h,w]; bin i = [min + i*interval, min + (i+1)*interval) evaluated in the
    tensor's dtype (0-dim tensors for the per-image range, python floats for the fixed [0,1] range)."""
    d = torch.squeeze(depth_map)
    hi, lo = (1, 0) if fixed_range else (torch.max(d), torch.min(d))
    interval = (hi - lo) / num
    planes = []
    for i in range(num):
        m = torch.zeros(d.shape)
        m[(d >= lo + interval * i) & (d < lo + interval * (i + 1))] = 1
        planes.append(m)
    return torch.stack(planes, 0)
