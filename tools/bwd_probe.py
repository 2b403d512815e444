"""Stage-T backward kernel on the res-256 layer shape (B=32, C=128, k=16): timing, and a target for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gansformer_b200 as gf
dev = torch.device("cuda:0")
B, res, C, k = int(os.environ.get("BP_BATCH", 32)), int(os.environ.get("BP_RES", 256)), int(os.environ.get("BP_C", 128)), 16
attn = gf.BipartiteAttention(C, 32, k).to(dev)
x = torch.randn(B, res, res, C, device=dev, requires_grad=True)
y = torch.randn(B, k, 32, device=dev, requires_grad=True)
g = torch.randn(B, res, res, C, device=dev)
for _ in range(2):
    out, _, _ = attn(x, y)
    out.backward(g)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
out, _, _ = attn(x, y)
e[1].record()
out.backward(g)
e[2].record()
torch.cuda.synchronize()
nb = 4 * B * res * res * C
print(f"res={res} C={C} B={B}: forward {e[0].elapsed_time(e[1]):.3f} ms, backward (kernel + 2 bmm + table autograd) {e[1].elapsed_time(e[2]):.3f} ms; "
      f"backward kernel algorithmic bytes (x, dOut read; dX, dCtl written) = {4 * nb / 1e9:.2f} GB")
