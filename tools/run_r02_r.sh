mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "native or torgb or generator or config or ops or graph" > gpurun_out/pytest_gpu_r.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r.log; tail -6 gpurun_out/pytest_gpu_r.log
python bench.py --no-cpu-baseline --no-train-probe --no-duplex-probe > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; tail -c 300 gpurun_out/bench_r.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["stage_T"]["frac"], d.get("value_cudnn_convs", {}).get("value"), d["gpu_launches"])
PY
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/step_launches_r.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-probe --no-duplex-probe --no-fp32-convs > /dev/null 2>&1
