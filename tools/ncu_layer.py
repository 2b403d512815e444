"""One attention layer call between cudaProfilerStart/Stop, for `ncu --profile-from-start off --set full` captures of the dominant
kernels as the generator runs them:  NL_RES (256), NL_C (128), NL_K (16), NL_B (32), NL_DUPLEX (0), NL_POST (1: fused in_scale /
noise / bias / lrelu / post_scale; 2: also the fused tRGB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gansformer_b200 as gf
dev = torch.device("cuda:0")
res, C, k, B = (int(os.environ.get(n, d)) for n, d in (("NL_RES", 256), ("NL_C", 128), ("NL_K", 16), ("NL_B", 32)))
duplex, post_mode = bool(int(os.environ.get("NL_DUPLEX", "0"))), int(os.environ.get("NL_POST", "1"))
torch.manual_seed(0)
x = torch.randn(B, res, res, C, device=dev)
y = torch.randn(B, k, 32, device=dev)
out = torch.empty_like(x)
attn = gf.BipartiteAttention(C, 32, k, kmeans=duplex).to(dev)
post = None
if post_mode:
    post = dict(bias=torch.randn(C, device=dev), noise=torch.randn(res, res, device=dev), strength=torch.tensor(0.1, device=dev), act="lrelu", gain=2 ** 0.5,
                in_scale=torch.rand(B, C, device=dev) + 0.5, post_scale=torch.rand(B, C, device=dev) + 0.5)
    if post_mode == 2:
        post.update(rgb_w=torch.randn(B, 3, C, device=dev) / C ** 0.5, rgb_bias=torch.zeros(3, device=dev), rgb_out=torch.empty(B, 3, res, res, device=dev))
with torch.no_grad():
    for _ in range(2):
        attn(x, y, out=out, postop=post, need_centroids=False)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    attn(x, y, out=out, postop=post, need_centroids=False)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done", gf._lib.last_path(), gf._lib.last_centroid_path() if duplex else "-")
