mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "torgb or generator or config" > gpurun_out/pytest_gpu_p.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_p.log; tail -6 gpurun_out/pytest_gpu_p.log
python bench.py --no-cpu-baseline --no-train-probe --no-duplex-probe --no-fp32-convs > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; tail -c 300 gpurun_out/bench_p.err
GF_NO_TORGB_EPILOGUE=1 python bench.py --no-cpu-baseline --no-train-probe --no-duplex-probe --no-fp32-convs > gpurun_out/bench_p_nofuse.json 2> /dev/null
python - <<'PY'
import json
for f in ("bench_p", "bench_p_nofuse"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["stage_T"]["frac"], d.get("value_cudnn_convs", {}).get("value"), d["gpu_launches"])
PY
