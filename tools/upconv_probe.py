"""cuDNN stride-2 transposed 3x3 convolution vs its 4-phase decomposition into stride-1 convolutions (2x2, 2x1, 1x2, 1x1
taps) on the low-resolution input: same flops, but the phases run cuDNN's fprop kernels instead of strided dgrad."""
import torch, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = True; torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
B = 32
for (H, Cin, Cout) in [(32, 512, 512), (64, 512, 256), (128, 256, 128), (16, 512, 512)]:
    x = torch.randn(B, Cin, H, H, device=dev).contiguous(memory_format=torch.channels_last)
    w = torch.randn(Cout, Cin, 3, 3, device=dev) / (Cin * 9) ** 0.5
    wT = w.transpose(0, 1).contiguous(memory_format=torch.channels_last)          # [Cin, Cout, 3, 3] for conv_transpose2d
    ph = {}
    for a in (0, 1):
        for b in (0, 1):
            ky = [2, 0] if a == 0 else [1]
            kx = [2, 0] if b == 0 else [1]
            ph[(a, b)] = (w[:, :, ky][:, :, :, kx].contiguous(memory_format=torch.channels_last), (1 if a == 0 else 0, 1 if b == 0 else 0))
    def ref(): return F.conv_transpose2d(x, wT, stride=2)
    def phases(): return [F.conv2d(x, wk, padding=pad) for (wk, pad) in ph.values()]
    with torch.no_grad():
        T = ref(); P = phases()
        err = 0.0
        for (a, b), p in zip(ph.keys(), P):
            err = max(err, (T[:, :, a::2, b::2] - p).abs().max().item())
        t_ref, t_ph = timeit(ref), timeit(phases)
    gf = 2 * B * H * H * 9 * Cin * Cout / 1e9
    print(f"H={H} {Cin}->{Cout}: conv_transpose {t_ref:.3f} ms ({gf / t_ref:.0f} TF/s)  4 phases {t_ph:.3f} ms ({gf / t_ph:.0f} TF/s)  max|diff|={err:.2e}")
