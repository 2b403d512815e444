mkdir -p gpurun_out
for i in 1 2 3; do
python bench.py --no-cpu-baseline --no-train-probe --no-duplex-probe --no-fp32-convs > gpurun_out/bench_q_a$i.json 2> gpurun_out/bench_q.err || tail -5 gpurun_out/bench_q.err
GF_TORGB_EPILOGUE_C256=1 python bench.py --no-cpu-baseline --no-train-probe --no-duplex-probe --no-fp32-convs > gpurun_out/bench_q_b$i.json 2> /dev/null
done
python - <<'PY'
import json
for f in ("a1","b1","a2","b2","a3","b3"):
    d = json.loads(open(f"gpurun_out/bench_q_{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), round(d["roofline"]["stage_T"]["frac"],3), d["gpu_launches"])
PY
