"""The 6 distinct duplex attention layer calls of BASELINE configs[2] (256^2, K = 32, batch 64 by default), each run once
between cudaProfilerStart/Stop after a warm-up.  Under
    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv
this gives the per-kernel time and DRAM bytes of every launch of the layer chain (pass A, merge, key products, finalize,
stage T).  Without ncu: prints CUDA-event time per layer call (graph replay, rotating inputs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gansformer_b200 as gf
dev = torch.device("cuda:0")
B = int(os.environ.get("DL_BATCH", 64)); k = int(os.environ.get("DL_K", 32))
duplex = bool(int(os.environ.get("DL_DUPLEX", "1")))
layers = [(8, 512), (16, 512), (32, 512), (64, 512), (128, 256), (256, 128)]
if os.environ.get("DL_ONLY"):
    layers = [l for l in layers if str(l[0]) in os.environ["DL_ONLY"].split(",")]
profile = bool(int(os.environ.get("DL_PROFILE", "0")))
peak = 6576.1
tot_ms = tot_b = 0.0
for res, C in layers:
    nb = 2 * 4 * B * res * res * C
    nbuf = 2 if nb // 2 > (200 << 20) else 4
    xs = [torch.randn(B, res, res, C, device=dev) for _ in range(nbuf)]
    y = torch.randn(B, k, 32, device=dev)
    out = torch.empty_like(xs[0])
    attn = gf.BipartiteAttention(C, 32, k, kmeans=duplex).to(dev)
    with torch.no_grad():
        for i in range(2):
            attn(xs[i % nbuf], y, out=out, need_centroids=False)
        torch.cuda.synchronize()
        if profile:
            torch.cuda.profiler.start()
            attn(xs[0], y, out=out, need_centroids=False)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            continue
        graphs = []
        for i in range(nbuf):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                attn(xs[i], y, out=out, need_centroids=False)
            graphs.append(g)
        for g in graphs:
            g.replay()
        reps = 3 * nbuf
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            graphs[i % nbuf].replay()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tot_ms += 2 * ms; tot_b += 2 * nb
    print(f"res={res:4d} C={C:4d} K={k} B={B} {'duplex' if duplex else 'simplex'} {ms:8.4f} ms  {nb / ms / 1e6:8.0f} GB/s  frac={nb / ms / 1e6 / peak:.3f}", flush=True)
    del xs, out, attn
    torch.cuda.empty_cache()
if not profile:
    print(f"SUM (each layer twice = 12 layers): {tot_ms:.3f} ms, {tot_b / 1e9:.3f} GB -> {tot_b / tot_ms / 1e6:.0f} GB/s, frac {tot_b / tot_ms / 1e6 / peak:.3f}")
