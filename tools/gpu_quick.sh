# quick GPU check of a change: the GPU tests selected by $1 (pytest -k expression), then the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "$1" > gpurun_out/pytest_gpu_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_quick.log; tail -6 gpurun_out/pytest_gpu_quick.log
python bench.py --no-cpu-baseline --no-duplex-probe --no-fp32-convs > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 300 gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["stage_T"]["frac"], d["gpu_launches"], d.get("train_step", {}).get("ms_per_step"))
PY
