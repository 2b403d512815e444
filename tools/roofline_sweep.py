"""BASELINE configs[4] "attention roofline sweep": achieved algorithmic bandwidth (2 * 4 * B * n * C bytes per layer call) of
the bipartite attention layer over (resolution, K latents, simplex / duplex), each call replayed from a CUDA graph, against the
measured HBM copy peak.  Batch: as many images as keep one activation tensor at <= 1 GiB (max 128)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gansformer_b200 as gf

dev = torch.device("cuda:0")
peak = 6576.1
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
layers = [(8, 512), (16, 512), (32, 512), (64, 512), (128, 256), (256, 128), (512, 64)]
print(f"{'res':>4} {'C':>4} {'K':>3} {'B':>4} {'mode':>8} {'ms':>8} {'GB/s':>8} {'frac':>6}  path")
for res, C in layers:
    B = int(min(128, max(1, (1 << 30) // (res * res * C * 4))))
    for k in (8, 16, 32):
        for duplex in (False, True):
            xs = [torch.randn(B, res, res, C, device=dev) for _ in range(2)]
            y = torch.randn(B, k, 32, device=dev)
            out = torch.empty_like(xs[0])
            attn = gf.BipartiteAttention(C, 32, k, kmeans=duplex).to(dev)
            with torch.no_grad():
                for i in range(2):
                    attn(xs[i], y, out=out, need_centroids=False)
                torch.cuda.synchronize()
                graphs = []
                for i in range(2):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        attn(xs[i], y, out=out, need_centroids=False)
                    graphs.append(g)
                for g in graphs:
                    g.replay()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(6):
                    graphs[i & 1].replay()
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 6
            nb = 2 * 4 * B * res * res * C
            print(f"{res:4d} {C:4d} {k:3d} {B:4d} {'duplex' if duplex else 'simplex':>8} {ms:8.4f} {nb / ms / 1e6:8.0f} {nb / ms / 1e6 / peak:6.3f}  "
                  f"{gf._lib.last_path()}" + (f" / pass A {gf._lib.last_centroid_path()}" if duplex else ""))
            del xs, out, attn, graphs
            torch.cuda.empty_cache()
