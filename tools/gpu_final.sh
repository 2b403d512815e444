# last GPU call of the round: the full GPU suite, then the default bench line (both under their own timeouts)
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
GF_PARITY_LOG=gpurun_out/parity_log.jsonl timeout 140 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log; tail -3 gpurun_out/pytest_gpu_final.log
timeout 120 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -c 200 gpurun_out/bench_final.err
