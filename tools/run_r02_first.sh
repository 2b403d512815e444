set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python tools/l2_probe.py > gpurun_out/l2_probe.log 2>&1
python tools/duplex_layers.py > gpurun_out/duplex_layers_base.log 2>&1
DL_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/duplex_layers_ncu_base.csv python tools/duplex_layers.py > gpurun_out/duplex_layers_ncu.log 2>&1
rm -f gpurun_out/parity_log.jsonl
GF_PARITY_LOG=gpurun_out/parity_log.jsonl timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err
bash tools/sanitize.sh gpurun_out
