set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_e.log
tail -4 gpurun_out/pytest_gpu_e.log
python bench.py --no-train-probe --no-cpu-baseline > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; tail -c 300 gpurun_out/bench_e.err
AB_POST=1 AB_MODES=default python tools/attn_bench.py > gpurun_out/attn_layers_post_e.log 2>&1; tail -8 gpurun_out/attn_layers_post_e.log
AB_POST=0 AB_MODES=default python tools/attn_bench.py > gpurun_out/attn_layers_plain_e.log 2>&1; tail -8 gpurun_out/attn_layers_plain_e.log
