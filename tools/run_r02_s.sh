mkdir -p gpurun_out
cat > /tmp/c256.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch, gansformer_b200 as gf
dev = torch.device("cuda:0")
for res, C, B, k in [(128, 256, 32, 16), (128, 256, 64, 16), (64, 256, 128, 16), (128, 256, 32, 32)]:
    xs = [torch.randn(B, res, res, C, device=dev) for _ in range(2)]
    y = torch.randn(B, k, 32, device=dev)
    out = torch.empty_like(xs[0])
    attn = gf.BipartiteAttention(C, 32, k).to(dev)
    post = dict(bias=torch.randn(C, device=dev), noise=torch.randn(res, res, device=dev), strength=torch.tensor(0.1, device=dev), act="lrelu", gain=1.414,
                in_scale=torch.rand(B, C, device=dev) + 0.5, post_scale=torch.rand(B, C, device=dev) + 0.5)
    for po in (None, post):
        with torch.no_grad():
            for i in range(2): attn(xs[i], y, out=out, need_centroids=False, postop=po)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(10): attn(xs[i & 1], y, out=out, need_centroids=False, postop=po)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        nb = 2 * 4 * B * res * res * C
        print(f"res {res} C {C} B {B} k {k} post {po is not None}: {ms:.4f} ms {nb / ms / 1e6:.0f} GB/s", flush=True)
PY
for i in 1 2; do
echo "== default"; python /tmp/c256.py
echo "== two-pass C256"; GF_ATTN_LIB=$PWD/gansformer-reproducibility-challenge_b200/libgf_attn_tp8.so python /tmp/c256.py
done
for i in 1 2; do
python bench.py --no-cpu-baseline --no-train-probe --no-duplex-probe --no-fp32-convs 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('default', d['value'], d['ms_per_step'], d['roofline']['frac'])"
GF_ATTN_LIB=$PWD/gansformer-reproducibility-challenge_b200/libgf_attn_tp8.so python bench.py --no-cpu-baseline --no-train-probe --no-duplex-probe --no-fp32-convs 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tp8', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
