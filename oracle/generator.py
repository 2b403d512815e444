"""CPU oracle of the full GANsformer generator forward (mapping + synthesis + attention), NCHW, direct op order.

TEST INFRASTRUCTURE ONLY; PARITY UNPINNED -- see oracle/bipartite.py.  The reference source (expected
src/training/network.py: G_GANsformer, G_mapping, G_synthesis, modulated_conv2d_layer; native ops
dnnlib/tflib/ops/{fused_bias_act,upfirdn_2d}.cu) is not in /root/reference, so this restates SURVEY.md
Appendix A + A.4 [SPEC] 1-9 instead of citing file:line.

The function consumes the *state_dict* of the product ``Generator`` (so both sides share bits of every
weight) but none of its code: the modulated convolution is done the way the reference would (per-sample
modulated + demodulated weights, grouped convolution), attention goes through
``oracle.bipartite.transformer_layer`` with its two transposes.  Runs in the dtype requested (float64 = truth,
float32 = "reference Python path" stand-in / timed CPU baseline).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .bipartite import transformer_layer

SQRT2 = math.sqrt(2.0)


def _fc(x, sd, prefix, in_features, lr_mul=1.0, act="linear"):
    w = sd[prefix + ".weight"] * (lr_mul / math.sqrt(in_features))
    b = sd[prefix + ".bias"] * lr_mul
    x = x @ w.t() + b
    return F.leaky_relu(x, 0.2) * SQRT2 if act == "lrelu" else x


def _fir(dtype):
    f = torch.tensor([1.0, 3.0, 3.0, 1.0], dtype=torch.float64)
    f = torch.outer(f, f)
    return (f / f.sum()).to(dtype)


def _upfirdn(x, f, up=1, pad=(0, 0, 0, 0), gain=1.0):
    B, C, H, W = x.shape
    if up > 1:
        z = torch.zeros(B, C, H * up, W * up, dtype=x.dtype)
        z[:, :, ::up, ::up] = x
        x = z
    x = F.pad(x, [pad[0], pad[1], pad[2], pad[3]])
    w = (f * gain)[None, None].expand(C, 1, *f.shape)
    return F.conv2d(x, w, groups=C)


def _modconv(x, weight, styles, demodulate=True, up=1, f=None):
    """Reference-style modulated conv: per-sample weights, grouped convolution."""
    B = x.shape[0]
    O, I, kh, kw = weight.shape
    w = weight[None] * (1.0 / math.sqrt(I * kh * kw)) * styles[:, None, :, None, None]       # [B,O,I,kh,kw]
    if demodulate:
        w = w * torch.rsqrt(w.square().sum(dim=[2, 3, 4], keepdim=True) + 1e-8)
    x = x.reshape(1, B * I, *x.shape[2:])
    if up == 1:
        x = F.conv2d(x, w.reshape(B * O, I, kh, kw), padding=kh // 2, groups=B)
    else:
        wt = w.transpose(1, 2).reshape(B * I, O, kh, kw)
        x = F.conv_transpose2d(x, wt, stride=2, groups=B)
        x = x.reshape(B, O, *x.shape[2:])
        return _upfirdn(x, f, pad=(1, 1, 1, 1), gain=4.0)
    return x.reshape(B, O, *x.shape[2:])


def generator_forward(sd: Dict[str, torch.Tensor], z: torch.Tensor, *, resolution: int, components_num: int,
                      latent_dim: int, integration="mul", norm="layer", duplex=False, use_pos=True, num_heads=1,
                      truncation_psi: float = 1.0, noise_mode: str = "const", mapping_layers: int = 8,
                      g_start_res: int = 8, g_end_res: Optional[int] = None, dtype=torch.float64,
                      return_att: bool = False, return_features: bool = False, kmeans_iters: int = 1, img2ltnt: bool = False,
                      iterative: bool = False):
    """z [B, k+1, D] -> img [B, 3, R, R] (NCHW).  `sd` = product Generator.state_dict() (any device/dtype)."""
    sd = {k_: v.detach().to("cpu", dtype) if v.is_floating_point() else v.detach().cpu() for k_, v in sd.items()}
    z = z.detach().to("cpu", dtype)
    k, D = components_num, latent_dim
    g_end_res = resolution if g_end_res is None else g_end_res
    B = z.shape[0]
    # ---- G_mapping
    z = z * torch.rsqrt(z.square().mean(dim=2, keepdim=True) + 1e-8)
    loc, glo = z[:, :k], z[:, k:]
    for i in range(mapping_layers):
        loc = _fc(loc, sd, f"mapping.local.{i}", D, lr_mul=0.01, act="lrelu")
        glo = _fc(glo, sd, f"mapping.glob.{i}", D, lr_mul=0.01, act="lrelu")
        pre = f"mapping.self_att.{i}."
        if (pre + "wq") in sd:      # ltnt2ltnt: the k local latents attend to each other after every layer:
            # the same block as in the synthesis network with the latents as the "image" [B, D, k, 1] and as the attended set
            w = {n[len(pre):]: t for n, t in sd.items() if n.startswith(pre)}
            xl, _, _ = transformer_layer(loc.transpose(1, 2).reshape(B, D, k, 1), loc, w, integration=integration, norm=norm,
                                         duplex=False, num_heads=num_heads, use_pos=False)
            loc = xl.reshape(B, D, k).transpose(1, 2)
    if truncation_psi != 1.0:
        loc = sd["mapping.w_avg"][0].lerp(loc, truncation_psi)
        glo = sd["mapping.w_avg"][1].lerp(glo, truncation_psi)
    y, w_glob = loc, glo[:, 0]
    # ---- G_synthesis
    f = _fir(dtype)
    x = sd["synthesis.const"][None].expand(B, -1, -1, -1)
    img = None
    atts: List[torch.Tensor] = []
    feats: List[torch.Tensor] = []
    li = 0
    cen_prev = None
    res_list = [2 ** i for i in range(2, int(math.log2(resolution)) + 1)]
    for bi, res in enumerate(res_list):
        for j in range(1 if res == 4 else 2):
            pre = f"synthesis.layers.{li}"
            li += 1
            up = 2 if (res > 4 and j == 0) else 1
            in_ch = sd[pre + ".weight"].shape[1]
            styles = _fc(w_glob, sd, pre + ".affine", D)
            x = _modconv(x, sd[pre + ".weight"], styles, up=up, f=f)
            if (pre + ".attention.wq") in sd and g_start_res <= res <= g_end_res:
                w = {n[len(pre) + 11:]: t for n, t in sd.items() if n.startswith(pre + ".attention.")}
                cen_init = cen_prev if (iterative and duplex and cen_prev is not None and cen_prev.shape[2] == x.shape[1]) else None
                x, att, cen_prev = transformer_layer(x, y, w, integration=integration, norm=norm, duplex=duplex,
                                                     num_heads=num_heads, use_pos=use_pos, return_att=return_att,
                                                     kmeans_iters=kmeans_iters, img2ltnt=img2ltnt, centroids_init=cen_init)
                if att is not None:
                    atts.append(att)
                has_att = True
            else:
                has_att = False
            if noise_mode == "const":
                x = x + sd[pre + ".noise_const"] * sd[pre + ".noise_strength"]
            elif noise_mode != "none":
                raise ValueError("the oracle supports noise_mode 'const' or 'none' (random noise is not reproducible)")
            x = F.leaky_relu(x + sd[pre + ".bias"][None, :, None, None], 0.2) * SQRT2
            if return_features and has_att:           # the layer's activation: attention -> noise -> bias -> leaky-ReLU
                feats.append(x)
            del in_ch
        pre = f"synthesis.torgbs.{bi}"
        styles = _fc(w_glob, sd, pre + ".affine", D)
        rgb = _modconv(x, sd[pre + ".weight"], styles, demodulate=False) + sd[pre + ".bias"][None, :, None, None]
        img = rgb if img is None else _upfirdn(img, f, up=2, pad=(2, 1, 2, 1), gain=4.0) + rgb
    out = (img,)
    if return_att:
        out += (atts,)
    if return_features:
        out += (feats,)
    return out[0] if len(out) == 1 else out
