"""One eager forward of the duplex generator (BASELINE configs[2]: 256x256, K=32, batch 64) between cudaProfilerStart/Stop:
run under `ncu --profile-from-start off --metrics gpu__time_duration.sum` to get the launch list of the step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gansformer_b200 as gf
torch.backends.cudnn.allow_tf32 = True; torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
torch.manual_seed(0)
G = gf.Generator(resolution=256, components_num=32, latent_dim=32, kmeans=True).to(dev).eval()
z = torch.randn(64, 33, 32, device=dev)
with torch.no_grad():
    for _ in range(2):
        G(z)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    G(z)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
