"""Where does the G/D training step spend its GPU time?  (torch.profiler, top CUDA kernels of one step at 256x256, B=16)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gansformer_b200 as gf
from importlib import import_module
tr = import_module("gansformer-reproducibility-challenge_b200.training")
torch.backends.cudnn.allow_tf32 = True; torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
B = int(os.environ.get("TP_BATCH", 16))
torch.manual_seed(0)
G = gf.Generator(resolution=256, components_num=16, latent_size=512).to(dev)
D = tr.Discriminator(256).to(dev)
trainer = tr.Trainer(G, D)
z = torch.randn(B, 17, G.latent_dim, device=dev)
reals = torch.rand(B, 3, 256, 256, device=dev) * 2 - 1
for _ in range(2):
    trainer.step(z, reals)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    trainer.step(z, reals)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
