set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
GF_PARITY_LOG=gpurun_out/parity_log.jsonl timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_c.log
tail -5 gpurun_out/pytest_gpu_c.log
python tools/duplex_layers.py > gpurun_out/duplex_layers_c.log 2>&1; cat gpurun_out/duplex_layers_c.log
python bench.py > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; tail -c 400 gpurun_out/bench_c.err
bash tools/traffic_capture.sh 2 gpurun_out
for c in 1 3 5; do python bench.py --config $c --steps 5 > gpurun_out/bench_c_config$c.json 2> gpurun_out/bench_c_config$c.err; tail -c 300 gpurun_out/bench_c_config$c.err; done
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/step_launches_c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-probe --no-duplex-probe --no-fp32-convs > /dev/null 2>&1
