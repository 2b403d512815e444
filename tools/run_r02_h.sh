set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_h.log
tail -4 gpurun_out/pytest_gpu_h.log
python tools/duplex_layers.py > gpurun_out/duplex_layers_h.log 2>&1; cat gpurun_out/duplex_layers_h.log
AB_POST=1 AB_MODES=default python tools/attn_bench.py > gpurun_out/attn_layers_post_h.log 2>&1; tail -7 gpurun_out/attn_layers_post_h.log
AB_POST=0 AB_MODES=default python tools/attn_bench.py > gpurun_out/attn_layers_plain_h.log 2>&1; tail -7 gpurun_out/attn_layers_plain_h.log
python bench.py --no-train-probe --no-cpu-baseline --no-duplex-probe --no-fp32-convs > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; tail -c 300 gpurun_out/bench_h.err
