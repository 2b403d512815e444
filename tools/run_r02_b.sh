set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_b.log
tail -8 gpurun_out/pytest_gpu_b.log
python tools/duplex_layers.py > gpurun_out/duplex_layers_b.log 2>&1; cat gpurun_out/duplex_layers_b.log
GF_CEN_LEAD=1000 python tools/duplex_layers.py > gpurun_out/duplex_layers_b_nolead.log 2>&1; cat gpurun_out/duplex_layers_b_nolead.log
GF_CEN_LEAD=3 DL_ONLY=128,256 python tools/duplex_layers.py; GF_CEN_LEAD=4 DL_ONLY=128,256 python tools/duplex_layers.py
DL_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/duplex_layers_ncu_b.csv python tools/duplex_layers.py > /dev/null 2>&1
python bench.py --no-cpu-baseline --no-train-probe > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -c 600 gpurun_out/bench_b.err
