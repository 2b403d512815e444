"""Training step: GPU time of the eager elementwise torch ops by input shape (which of them are worth a native fused op)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gansformer_b200 as gf
from importlib import import_module
tr = import_module("gansformer-reproducibility-challenge_b200.training")
torch.backends.cudnn.allow_tf32 = True; torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
B = int(os.environ.get("TP_BATCH", 32))
torch.manual_seed(0)
G = gf.Generator(resolution=256, components_num=16, latent_size=512).to(dev)
D = tr.Discriminator(256).to(dev)
trainer = tr.Trainer(G, D)
z = torch.randn(B, 17, G.latent_dim, device=dev)
reals = torch.rand(B, 3, 256, 256, device=dev) * 2 - 1
trainer.it = 1
for _ in range(2):
    trainer.step(z, reals); trainer.it = 1
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=False) as prof:
    trainer.step(z, reals)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    t = getattr(e, "self_device_time_total", 0) or getattr(e, "self_cuda_time_total", 0)
    if t > 300 and e.key.startswith("aten::") and not e.key.startswith("aten::conv") and not e.key.startswith("aten::cudnn") and "convolution" not in e.key:
        rows.append((t, e.key, e.count, str(e.input_shapes)[:110]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"eager non-conv aten ops with > 0.3 ms: {tot/1e3:.1f} ms")
for t, k, c, sh in rows[:45]:
    print(f"{t/1e3:8.2f} ms  x{c:3d}  {k:28s} {sh}")
