"""Compare the raw pass-A partials (acc, m, l) of the tcgen05 centroid kernel with the CUDA-core kernel."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gansformer_b200 as gf
from oracle import bipartite as ob, folded as of
dev = torch.device("cuda:0")
L = gf._lib
for (C, H, W, k, B) in [(64, 8, 16, 4, 1), (128, 16, 16, 16, 2)]:
    D = p = 32
    g = torch.Generator().manual_seed(C + k)
    x = torch.randn(B, C, H, W, generator=g, dtype=torch.float64) * 1.3 + 0.2
    y = torch.randn(B, k, D, generator=g, dtype=torch.float64)
    w = ob.init_params(C, D, k, p, "mul", True, seed=7, bias_std=0.4)
    # oracle pass A internals
    f = of.fold_weights(w, C=C, k=k, integration="mul", duplex=True)
    X = x.permute(0, 2, 3, 1).reshape(B, H * W, C)
    cen_ref, xbar_ref = of.centroid_pass(X, y, f, H=H, W=W, p=p)
    res = {}
    for exact in (False, True):
        attn = gf.BipartiteAttention(C, D, k, pos_dim=p, kmeans=True, exact_fp32=exact).to(dev)
        with torch.no_grad():
            for n_, prm in attn.named_parameters():
                prm.copy_(w[n_].float())
            out, _, cen = attn(x.permute(0, 2, 3, 1).contiguous().float().to(dev), y.float().to(dev))
            torch.cuda.synchronize()
        desc = L.make_desc(B, H, W, C, k, D, pos_dim=p, duplex=True, flags=1 if exact else 0)
        o = (ctypes.c_longlong * 8)()
        L.check(L.load().gf_attn_debug_layout(ctypes.byref(desc), o, 8), "dbg")
        wpart, wxbar, nsplit, KP = o[0], o[1], o[2], o[3]
        ws = list(attn._plan.ws.values())[0].view(torch.float32)
        part = ws[wpart:wpart + B * nsplit * KP * (C + 4)].reshape(B, nsplit, KP, C + 4).double().cpu()
        xbar = ws[wxbar:wxbar + B * k * C].reshape(B, k, C).double().cpu()
        res[exact] = (part, xbar, L.last_centroid_path())
    pt, xb_t, path_t = res[False]
    ps, xb_s, path_s = res[True]
    print(f"== C={C} {H}x{W} k={k} B={B}: paths tc={path_t} simt={path_s} nsplit={pt.shape[1]} KP={pt.shape[2]}")
    print("xbar err simt vs ref:", (xb_s - xbar_ref).abs().max().item(), " tc vs ref:", (xb_t - xbar_ref).abs().max().item())
    for j in range(min(k, 3)):
        for name, pp in (("tc", pt), ("simt", ps)):
            m, l = pp[0, 0, j, C].item(), pp[0, 0, j, C + 1].item()
            acc = pp[0, 0, j, :C]
            print(f" latent {j} {name:4s}: m={m:.4f} l={l:.4f} log(sum e^L)={m + (torch.log(torch.tensor(l)).item() if l > 0 else float('nan')):.4f} "
                  f"acc/l[:5]={(acc[:5] / l).tolist()} ")
        print("        ref xbar[:5] =", xbar_ref[0, j, :5].tolist())
    print(" tc   rows >= k (should be m=-inf,l=0):", pt[0, 0, k:min(k + 2, pt.shape[2]), C:C + 2].tolist())
