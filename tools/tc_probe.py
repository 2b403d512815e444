"""Bring-up probe for the tcgen05 stage-T kernel: one layer per shape, TC path vs fp64 oracle.
att error localises GEMM1/TMA/descriptors; out error adds GEMM2 (A from TMEM), LayerNorm and the epilogue."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gansformer_b200 as gf
from oracle import bipartite as ob

dev = torch.device("cuda:0")
shapes = [  # C, H, W, k, integration, norm, B
    (64, 8, 16, 4, "mul", "layer", 1),
    (64, 8, 16, 4, "mul", "none", 1),
    (128, 16, 16, 16, "mul", "layer", 2),
    (128, 32, 32, 16, "both", "layer", 3),
    (256, 16, 16, 16, "mul", "layer", 2),
    (128, 16, 16, 32, "add", "layer", 2),
    (256, 32, 32, 32, "mul", "layer", 5),
    (64, 64, 64, 8, "both", "layer", 40),
    (128, 128, 128, 16, "mul", "layer", 20),
    (512, 16, 16, 16, "mul", "layer", 2),
    (512, 32, 32, 8, "both", "layer", 3),
    (512, 16, 16, 32, "mul", "none", 2),
    (512, 64, 64, 16, "mul", "layer", 10),
    (256, 64, 64, 32, "both", "layer", 6),
]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
for (C, H, W, k, integ, norm, B) in shapes:
    D = p = 32
    g = torch.Generator().manual_seed(C + k)
    x = torch.randn(B, C, H, W, generator=g, dtype=torch.float64) * 1.3 + 0.2
    y = torch.randn(B, k, D, generator=g, dtype=torch.float64)
    w = ob.init_params(C, D, k, p, integ, False, seed=7, bias_std=0.4)
    nrm = None if norm == "none" else norm
    ref, ratt, _ = ob.transformer_layer(x, y, w, integration=integ, norm=nrm, return_att=True)
    ref = ref.permute(0, 2, 3, 1)
    attn = gf.BipartiteAttention(C, D, k, pos_dim=p, integration=integ, norm=nrm).to(dev)
    with torch.no_grad():
        for n_, prm in attn.named_parameters():
            prm.copy_(w[n_].float())
        xin = x.permute(0, 2, 3, 1).contiguous().float().to(dev)
        out, att, _ = attn(xin, y.float().to(dev), return_att=True)
        torch.cuda.synchronize()
    path = gf._lib.last_path()
    e = (out.double().cpu() - ref).abs()
    ea = (att.double().cpu() - ratt).abs()
    rel = (e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    ratio = (e / (8e-3 + 8e-3 * ref.abs())).max().item()
    # where are the errors? per-image / per-128-token tile / per-32-channel slab maxima
    en = e.reshape(B, H * W, C)
    per_img = en.amax(dim=(1, 2))
    per_slab = en.reshape(B, H * W, C // 32, 32).amax(dim=(0, 1, 3))
    print(f"C={C} {H}x{W} k={k} {integ} {norm} B={B} path={path}: out max_abs={e.max():.3e} rel_rms={rel:.3e} tol_ratio={ratio:.2f} "
          f"att max_abs={ea.max():.3e} finite={bool(torch.isfinite(out).all())}")
    if ratio > 1 or not torch.isfinite(out).all():
        print("   per-image max:", [f"{v:.2e}" for v in per_img.tolist()][:8])
        print("   per-slab  max:", [f"{v:.2e}" for v in per_slab.tolist()])
        t = en[0].amax(dim=1).reshape(-1, 128).amax(dim=1)
        print("   per-tile (img 0) max:", [f"{v:.2e}" for v in t.tolist()][:8])
        r = en[0, :128].amax(dim=1)
        print("   per-row (img0 tile0) max, rows 0..15:", [f"{v:.2e}" for v in r[:16].tolist()])
        print("   sample got/ref row0 ch0..7:", out[0, 0, 0, :8].tolist(), ref[0, 0, 0, :8].tolist())
