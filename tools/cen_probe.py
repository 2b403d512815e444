"""Bring-up probe for the tcgen05 duplex pass-A kernel: centroids vs the fp64 oracle on a few shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gansformer_b200 as gf
from oracle import bipartite as ob
dev = torch.device("cuda:0")
shapes = [(64, 8, 16, 4, 1), (128, 16, 16, 16, 2), (128, 32, 32, 32, 3), (256, 32, 16, 16, 2), (64, 64, 64, 8, 5), (256, 64, 64, 32, 4), (128, 128, 128, 16, 3), (512, 16, 16, 16, 2), (512, 32, 32, 32, 3), (512, 64, 64, 8, 9)]
for (C, H, W, k, B) in shapes:
    D = p = 32
    g = torch.Generator().manual_seed(C + k)
    x = torch.randn(B, C, H, W, generator=g, dtype=torch.float64) * 1.3 + 0.2
    y = torch.randn(B, k, D, generator=g, dtype=torch.float64)
    w = ob.init_params(C, D, k, p, "mul", True, seed=7, bias_std=0.4)
    ref, _, rcen = ob.transformer_layer(x, y, w, duplex=True)
    attn = gf.BipartiteAttention(C, D, k, pos_dim=p, kmeans=True).to(dev)
    with torch.no_grad():
        for n_, prm in attn.named_parameters():
            prm.copy_(w[n_].float())
        out, _, cen = attn(x.permute(0, 2, 3, 1).contiguous().float().to(dev), y.float().to(dev))
        torch.cuda.synchronize()
    e = (cen.double().cpu() - rcen).abs()
    eo = (out.double().cpu() - ref.permute(0, 2, 3, 1)).abs()
    rel = (e.pow(2).mean().sqrt() / rcen.pow(2).mean().sqrt()).item()
    print(f"C={C} {H}x{W} k={k} B={B} cen_path={gf._lib.last_centroid_path()} path={gf._lib.last_path()}: cen max_abs={e.max():.3e} rel_rms={rel:.3e} "
          f"ratio={(e / (8e-3 + 8e-3 * rcen.abs())).max():.2f} | out max_abs={eo.max():.3e} finite={bool(torch.isfinite(cen).all())}")
    if not torch.isfinite(cen).all() or rel > 5e-3:
        print("   per-latent max err (img 0):", [f"{v:.2e}" for v in e[0].amax(dim=1).tolist()])
        print("   per-32ch-slab max err:", [f"{v:.2e}" for v in e.reshape(B, k, C // 32, 32).amax(dim=(0, 1, 3)).tolist()])
        print("   got[0,0,:6]", cen[0, 0, :6].tolist(), " ref", rcen[0, 0, :6].tolist())
