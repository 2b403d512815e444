set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
GF_PARITY_LOG=gpurun_out/parity_log.jsonl timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_d.log
tail -25 gpurun_out/pytest_gpu_d.log
python bench.py --no-train-probe > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; tail -c 400 gpurun_out/bench_d.err
GF_NO_TORGB_EPILOGUE=1 python bench.py --no-train-probe --no-cpu-baseline --no-duplex-probe --no-fp32-convs > gpurun_out/bench_d_notorgb.json 2>/dev/null
GF_NO_MAPPING_KERNEL=1 python bench.py --no-train-probe --no-cpu-baseline --no-duplex-probe --no-fp32-convs > gpurun_out/bench_d_nomap.json 2>/dev/null
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/step_launches_d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-probe --no-duplex-probe --no-fp32-convs > /dev/null 2>&1
