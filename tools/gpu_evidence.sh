mkdir -p gpurun_out
SAN_TIMEOUT=500 bash tools/sanitize.sh gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/step_launches_n.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-probe --no-duplex-probe --no-fp32-convs > /dev/null 2>&1
cat > /tmp/conv_one.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, gansformer_b200 as gf
from importlib import import_module
ops = import_module("gansformer-reproducibility-challenge_b200.ops")
dev = torch.device("cuda:0")
res, ci, co = int(os.environ.get("R", 128)), int(os.environ.get("CI", 256)), int(os.environ.get("CO", 256))
x = torch.randn(32, ci, res, res, device=dev).contiguous(memory_format=torch.channels_last)
wt = ops.conv3x3_pack(torch.randn(co, ci, 3, 3, device=dev) / (ci * 9) ** 0.5)
for _ in range(2): ops.conv3x3_native(x, wt)
torch.cuda.synchronize(); torch.cuda.profiler.start(); ops.conv3x3_native(x, wt); torch.cuda.synchronize(); torch.cuda.profiler.stop()
PY
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv3x3 -o gpurun_out/ncu_full_conv3x3_v2_res128_C256 python /tmp/conv_one.py > /dev/null 2>&1
R=256 CI=128 CO=128 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv3x3 -o gpurun_out/ncu_full_conv3x3_v2_res256_C128 python /tmp/conv_one.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
