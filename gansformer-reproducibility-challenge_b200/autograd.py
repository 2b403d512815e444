"""Differentiable wrapper of the fused attention op (needed by the G/D training step).

Forward = the C-ABI CUDA kernels.  Backward = PyTorch autograd through a recomputation of the same folded
algebra with torch ops on the GPU ("first via PyTorch autograd on the oracle-equivalent
composite"); a hand-written backward kernel is the follow-up.
"""
from __future__ import annotations

import math

import torch


def _e(w):
    return w * (1.0 / math.sqrt(w.shape[0]))


def _axis(length, dim, device):
    pos = (torch.arange(length, dtype=torch.float64, device=device) + 0.5) / length * 2.0 - 1.0
    freq = (math.pi / 2.0) * torch.pow(2.0, torch.arange(dim // 2, dtype=torch.float64, device=device))
    ang = pos[:, None] * freq[None, :]
    return torch.cat([torch.sin(ang), torch.cos(ang)], dim=1).float()


def composite_forward(x, y, p, *, integration, norm, duplex, use_pos, centroids=None):
    """Same math as the kernels, in torch ops (direct op order), on whatever device x lives on.  x [B,H,W,C]."""
    B, H, W, C = x.shape
    n = H * W
    X = x.reshape(B, n, C)
    s = 1.0 / math.sqrt(C)
    if use_pos:
        pd = p["pos_latent"].shape[1]
        half = pd // 2
        row, col = _axis(H, half, x.device), _axis(W, half, x.device)
        Pg = torch.cat([row[:, None, :].expand(H, W, half), col[None, :, :].expand(H, W, half)], dim=2).reshape(n, pd)
        Pl = p["pos_latent"]
    cen = None
    if duplex:
        if centroids is not None:
            cen = centroids
        else:
            Qy = y @ _e(p["wq2"]) + p["bq2"]
            Kx = X @ _e(p["wk2"]) + p["bk2"]
            if use_pos:
                Qy = Qy + (Pl @ _e(p["wpq2"]))[None]
                Kx = Kx + (Pg @ _e(p["wpk2"]))[None]
            Vx = X @ _e(p["wv2"]) + p["bv2"]
            A = torch.softmax((Qy @ Kx.transpose(1, 2)) * s, dim=2)
            cen = A @ Vx
        K = cen @ _e(p["wkc"]) + p["bk"]
    else:
        K = y @ _e(p["wk"]) + p["bk"]
    Q = X @ _e(p["wq"]) + p["bq"]
    if use_pos:
        K = K + (Pl @ _e(p["wpk"]))[None]
        Q = Q + (Pg @ _e(p["wpq"]))[None]
    V = y @ _e(p["wv"]) + p["bv"]
    P = torch.softmax((Q @ K.transpose(1, 2)) * s, dim=2)
    ctl = (P @ V) @ _e(p["wo"]) + p["bo"]
    if norm == "layer":
        mu = X.mean(dim=2, keepdim=True)
        Xn = (X - mu) * torch.rsqrt(((X - mu) ** 2).mean(dim=2, keepdim=True) + 1e-8)
    elif norm in (None, "none"):
        Xn = X
    else:
        dims = (1,) if norm == "instance" else (0, 1)
        mu = X.mean(dim=dims, keepdim=True)
        Xn = (X - mu) * torch.rsqrt(((X - mu) ** 2).mean(dim=dims, keepdim=True) + 1e-8)
    if integration == "mul":
        out = Xn * (1.0 + ctl)
    elif integration == "add":
        out = Xn + ctl
    else:
        out = Xn * (1.0 + ctl[..., :C]) + ctl[..., C:]
    return out.reshape(B, H, W, C), cen


class _FusedAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, centroids, return_att, names, x, y, *params):
        from .attention import bipartite_attention_forward
        pd = dict(zip(names, params))
        out, att, cen = bipartite_attention_forward(
            x.detach(), y.detach(), {k: v.detach() for k, v in pd.items()}, module._plan,
            integration=module.integration, norm=module.norm, duplex=module.duplex, num_heads=module.num_heads,
            use_pos=module.use_pos, return_att=return_att, centroids=centroids, exact_fp32=module.exact_fp32,
            weights_version=tuple((v.data_ptr(), v._version) for v in params))
        ctx.module, ctx.names, ctx.centroids = module, names, centroids
        ctx.save_for_backward(x, y, *params)
        ctx.mark_non_differentiable(*[t for t in (att, cen) if t is not None])
        return out, att, cen

    @staticmethod
    def backward(ctx, g_out, g_att, g_cen):
        m = ctx.module
        x, y, *params = ctx.saved_tensors
        with torch.enable_grad():
            xs = x.detach().requires_grad_(True)
            ys = y.detach().requires_grad_(True)
            ps = [p.detach().requires_grad_(True) for p in params]
            out, _ = composite_forward(xs, ys, dict(zip(ctx.names, ps)), integration=m.integration, norm=m.norm,
                                       duplex=m.duplex, use_pos=m.use_pos, centroids=ctx.centroids)
            grads = torch.autograd.grad(out, [xs, ys, *ps], g_out, allow_unused=True)
        return (None, None, None, None, *grads)


def bipartite_attention_autograd(module, x, y, centroids, return_att):
    names = tuple(n for n, _ in module.named_parameters(recurse=False))
    params = tuple(p for _, p in module.named_parameters(recurse=False))
    return _FusedAttention.apply(module, centroids, return_att, names, x, y, *params)
