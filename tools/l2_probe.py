"""How much of a just-read tensor does the B200's L2 (126 MB over two dies) still hold?  ("measure the
effective resident size with a read-after-write microbenchmark before relying on it".)

For S in a sweep: pass 1 streams a buffer of S bytes (x.sum(): what duplex pass A does), pass 2 = y = x * 2 (what stage T does:
read x again, write S bytes).  Reported: time of pass 2 alone and the HBM-equivalent bandwidth (2 S / t); when x is still in
L2, pass 2 only writes and its "bandwidth" exceeds the HBM peak.  A cold pass 2 (L2 flushed in between) is the control."""
import os, sys, json
import torch
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
out = []
print(f"{'MB':>6} {'warm us':>9} {'cold us':>9} {'warm GB/s':>10} {'cold GB/s':>10}")
for mb in (8, 16, 24, 32, 40, 48, 56, 64, 72, 80, 96, 112, 128, 160, 192, 256):
    n = mb * (1 << 20) // 4
    x = torch.randn(n, device=dev)
    y = torch.empty_like(x)
    tw, tc = [], []
    for rep in range(7):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        flush.zero_()
        s = x.sum()                      # pass 1: stream x
        e[0].record(); torch.mul(x, 2.0, out=y); e[1].record()      # pass 2 warm
        flush.zero_()
        e[2].record(); torch.mul(x, 2.0, out=y); e[3].record()      # pass 2 cold
        torch.cuda.synchronize()
        tw.append(e[0].elapsed_time(e[1])); tc.append(e[2].elapsed_time(e[3]))
    w, c = sorted(tw)[len(tw) // 2] * 1e3, sorted(tc)[len(tc) // 2] * 1e3
    print(f"{mb:6d} {w:9.1f} {c:9.1f} {2 * mb * 1.048576 / w * 1e3:10.0f} {2 * mb * 1.048576 / c * 1e3:10.0f}", flush=True)
    out.append(dict(mb=mb, warm_us=w, cold_us=c))
    del x, y
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/l2_probe.json", "w"), indent=1)
