mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
GF_PARITY_LOG=gpurun_out/parity_log.jsonl timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_o.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_o.log; tail -4 gpurun_out/pytest_gpu_o.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench_o.json 2> gpurun_out/bench_o.err; tail -c 300 gpurun_out/bench_o.err
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/step_launches_o.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-probe --no-duplex-probe --no-fp32-convs > /dev/null 2>&1
for c in 1 3 5; do python bench.py --config $c --no-train-probe > gpurun_out/bench_config$c.json 2> gpurun_out/bench_config$c.err || tail -3 gpurun_out/bench_config$c.err; tail -n1 gpurun_out/bench_config$c.json | cut -c1-200; done
