"""CPU oracle for the GANsformer bipartite (simplex / duplex) attention block.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` may be imported by the
product package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or
as the timed CPU baseline.

PARITY UNPINNED.  The reference checkout at /root/reference holds no source
(``/root/reference/.SUBMODULES.json:2`` reports ``"bytes": 0``; the only other
files are ``LICENSE`` and ``src/Dockerfile``) and TensorFlow 1.14
(``/root/reference/src/Dockerfile:7``) is not installable here, so no golden
vector owned by the reference exists for this path.  This file restates the
algorithm from SURVEY.md Appendix A (paper equations + recollected structure of
the upstream ``src/training/network.py``: ``transformer_layer``, ``integrate``,
``att_norm``, ``dense_layer``, ``get_positional_embeddings`` -- none of them on
disk, so no file:line can be cited) in the *direct, unfolded* op order the
reference's TensorFlow graph would execute: NCHW -> [B,n,C] transpose, three
dense projections, QK^T, softmax, PV, output dense, norm, modulate, transpose
back.  It is deliberately written with plain tensor ops, one per reference op.

All arithmetic runs in the dtype of the inputs (float64 = truth, float32 = the
stand-in for "the reference's own Python path").
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

Tensor = torch.Tensor

LN_EPS = 1e-8


# ----------------------------------------------------------------------------
# positional embeddings (a5: get_positional_embeddings)
# ----------------------------------------------------------------------------
def sinusoidal_axis(length: int, dim: int, dtype=torch.float64) -> Tensor:
    """1-D sinusoidal table [length, dim]: ``dim/2`` sines then ``dim/2`` cosines.

    [SPEC] position of cell i is its centre mapped to (-1, 1); frequency m is
    ``(pi/2) * 2**m``.
    """
    assert dim % 2 == 0
    pos = (torch.arange(length, dtype=torch.float64) + 0.5) / length * 2.0 - 1.0
    freq = (math.pi / 2.0) * torch.pow(2.0, torch.arange(dim // 2, dtype=torch.float64))
    ang = pos[:, None] * freq[None, :]
    return torch.cat([torch.sin(ang), torch.cos(ang)], dim=1).to(dtype)


def grid_pos_table(H: int, W: int, pos_dim: int, dtype=torch.float64) -> Tensor:
    """2-D grid table [H*W, pos_dim] = concat(row_emb[h], col_emb[w])."""
    assert pos_dim % 4 == 0
    half = pos_dim // 2
    row = sinusoidal_axis(H, half, dtype)  # [H, half]
    col = sinusoidal_axis(W, half, dtype)  # [W, half]
    tab = torch.cat([row[:, None, :].expand(H, W, half), col[None, :, :].expand(H, W, half)], dim=2)
    return tab.reshape(H * W, pos_dim)


# ----------------------------------------------------------------------------
# parameters (a4: dense_layer / get_weight with equalised learning rate)
# ----------------------------------------------------------------------------
SIMPLEX_KEYS = ("wq", "bq", "wpq", "wk", "bk", "wpk", "wv", "bv", "wo", "bo", "pos_latent")
DUPLEX_KEYS = ("wq2", "bq2", "wpq2", "wk2", "bk2", "wpk2", "wv2", "bv2", "wkc")


def param_shapes(C: int, D: int, k: int, pos_dim: int, integration: str, duplex: bool, extras: bool = False) -> Dict[str, Tuple[int, ...]]:
    """Raw (un-scaled) parameter shapes of one attention layer.  Weights are [fan_in, fan_out]."""
    cout = 2 * C if integration == "both" else C
    shapes = {
        "wq": (C, C), "bq": (C,), "wpq": (pos_dim, C),
        "wk": (D, C), "bk": (C,), "wpk": (pos_dim, C),
        "wv": (D, C), "bv": (C,),
        "wo": (C, cout), "bo": (cout,),
        "pos_latent": (k, pos_dim),
    }
    if duplex:
        shapes.update({
            "wq2": (D, C), "bq2": (C,), "wpq2": (pos_dim, C),
            "wk2": (C, C), "bk2": (C,), "wpk2": (pos_dim, C),
            "wv2": (C, C), "bv2": (C,),
            "wkc": (C, C),
        })
        if extras:      # kmeans_iters > 1: centroid -> query projection; g_img2ltnt: centroid -> latent gain
            shapes.update({"wcq": (C, C), "wi2l": (C, D), "bi2l": (D,)})
    return shapes


def init_params(C: int, D: int, k: int, pos_dim: int, integration: str = "mul", duplex: bool = False,
                seed: int = 0, dtype=torch.float64, bias_std: float = 0.0, extras: bool = False) -> Dict[str, Tensor]:
    """N(0,1) weights, biases N(0, bias_std) (0 in benchmarks, >0 in tests so every term is live)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shp in param_shapes(C, D, k, pos_dim, integration, duplex, extras).items():   # extras come last: earlier draws are unchanged
        t = torch.randn(shp, generator=g, dtype=torch.float64)
        if name.startswith("b"):
            t = t * bias_std
        out[name] = t.to(dtype)
    return out


def _dense(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    """Equalised-LR dense: x @ (w / sqrt(fan_in)) + b."""
    y = x @ (w * (1.0 / math.sqrt(w.shape[0])))
    return y if b is None else y + b


# ----------------------------------------------------------------------------
# a3: att_norm / integrate
# ----------------------------------------------------------------------------
def att_norm(x: Tensor, norm: Optional[str], eps: float = LN_EPS) -> Tensor:
    """x [B,n,C].  layer: per (b,token) over C; instance: per (b,c) over n; batch: per c over (b,n)."""
    if norm is None or norm == "none":
        return x
    dims = {"layer": (2,), "instance": (1,), "batch": (0, 1)}[norm]
    mu = x.mean(dim=dims, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=dims, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps)


def integrate(x: Tensor, control: Tensor, integration: str, norm: Optional[str]) -> Tensor:
    """x [B,n,C], control [B,n,C or 2C] (already through its dense).  SURVEY A.2 last three lines."""
    C = x.shape[-1]
    xh = att_norm(x, norm)
    if integration == "mul":
        return xh * (1.0 + control)
    if integration == "add":
        return xh + control
    if integration == "both":
        return xh * (1.0 + control[..., :C]) + control[..., C:]
    raise ValueError(integration)


# ----------------------------------------------------------------------------
# a2: transformer_layer
# ----------------------------------------------------------------------------
def _split_heads(t: Tensor, h: int) -> Tensor:
    B, L, C = t.shape
    return t.reshape(B, L, h, C // h).permute(0, 2, 1, 3)  # [B,h,L,C/h]


def transformer_layer(x_nchw: Tensor, y: Tensor, w: Dict[str, Tensor], *, integration: str = "mul",
                      norm: Optional[str] = "layer", duplex: bool = False, num_heads: int = 1,
                      use_pos: bool = True, return_att: bool = False,
                      centroids_in: Optional[Tensor] = None, kmeans_iters: int = 1, img2ltnt: bool = False,
                      centroids_init: Optional[Tensor] = None, att_mult: Optional[Tensor] = None):
    """Bipartite attention, direct form.

    x_nchw [B,C,H,W]; y [B,k,D] (the k local latents).  Returns (x' [B,C,H,W], att [B,k,H,W] or None,
    centroids [B,k,C] or None).

    kmeans_iters > 1 (duplex; SURVEY A.3 "repeat with Qy derived from Cen"): iteration i >= 2 takes its queries from the previous
    centroids, Qy = dense(Cen, wcq) + bq2 (+ latent positional term); keys / values of the grid are unchanged.
    img2ltnt (duplex; SURVEY A.3 [SPEC] g_img2ltnt): before pass B the latents are modulated by the centroids,
    Y <- LN(Y) (1 + dense(Cen, wi2l) + bi2l) (layer norm over D, eps as att_norm, no affine); the values of pass B come from
    the modulated latents, the keys from the centroids.  The update is local to the layer (the caller's latents are not changed).
    centroids_init (duplex; `iterative=True` upstream, [SPEC]): centroids carried over from the previous attention layer of the same
    channel width initialise the k-means: the FIRST iteration already takes its queries from them (through wcq) instead of from the
    latents.
    att_mult [B,n,k] (training; att_dp): dropout multipliers (0 or 1/(1-p), oracle/philox.py) applied to the probabilities before they
    weight the values; the returned attention map is the probabilities before dropout.
    """
    B, C, H, W = x_nchw.shape
    n = H * W
    k = y.shape[1]
    dt = x_nchw.dtype
    h = num_heads
    scale = 1.0 / math.sqrt(C / h)

    X = x_nchw.reshape(B, C, n).permute(0, 2, 1)  # the reference's NCHW -> [B,n,C] transpose (a1)
    pos_dim = w["pos_latent"].shape[1]
    Pg = grid_pos_table(H, W, pos_dim, dt) if use_pos else None
    Pl = w["pos_latent"] if use_pos else None

    centroids = None
    if duplex:
        # pass A: latents attend to the image, softmax over the n grid cells
        if centroids_in is not None:
            centroids = centroids_in
        else:
            Qy = _dense(y, w["wq2"], w["bq2"])
            Kx = _dense(X, w["wk2"], w["bk2"])
            if use_pos:
                Qy = Qy + _dense(Pl, w["wpq2"])[None]
                Kx = Kx + _dense(Pg, w["wpk2"])[None]
            Vx = _dense(X, w["wv2"], w["bv2"])
            centroids = centroids_init
            for it in range(max(1, kmeans_iters)):
                if it > 0 or centroids_init is not None:     # queries from the previous centroids (carried in, or of the last iteration)
                    Qy = _dense(centroids, w["wcq"], w["bq2"])
                    if use_pos:
                        Qy = Qy + _dense(Pl, w["wpq2"])[None]
                A = torch.softmax((Qy @ Kx.transpose(1, 2)) * (1.0 / math.sqrt(C)), dim=2)  # [B,k,n] over n
                centroids = A @ Vx  # [B,k,C]
        K = _dense(centroids, w["wkc"], w["bk"])
    else:
        K = _dense(y, w["wk"], w["bk"])
    if use_pos:
        K = K + _dense(Pl, w["wpk"])[None]

    Q = _dense(X, w["wq"], w["bq"])
    if use_pos:
        Q = Q + _dense(Pg, w["wpq"])[None]
    yv = y
    if duplex and img2ltnt:
        yv = att_norm(y, "layer") * (1.0 + _dense(centroids, w["wi2l"], w["bi2l"]))
    V = _dense(yv, w["wv"], w["bv"])

    Qh, Kh, Vh = _split_heads(Q, h), _split_heads(K, h), _split_heads(V, h)
    S = (Qh @ Kh.transpose(2, 3)) * scale          # [B,h,n,k]
    P = torch.softmax(S, dim=3)                    # over the k latents
    Pd = P if att_mult is None else P * att_mult.to(dt)[:, None]      # attention dropout (same mask for every head: single-head layers only)
    ctrl = (Pd @ Vh).permute(0, 2, 1, 3).reshape(B, n, C)
    control = _dense(ctrl, w["wo"], w["bo"])       # gain (| bias)
    Xo = integrate(X, control, integration, norm)

    x_out = Xo.permute(0, 2, 1).reshape(B, C, H, W)  # transpose back (a1)
    att = None
    if return_att:
        att = P.mean(dim=1).permute(0, 2, 1).reshape(B, k, H, W)  # a6
    return x_out, att, centroids
