mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
GF_PARITY_LOG=gpurun_out/parity_log.jsonl timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_k.log; tail -6 gpurun_out/pytest_gpu_k.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; tail -c 500 gpurun_out/bench_k.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_k_ref.json 2>/dev/null; cut -c1-200 gpurun_out/bench_k_ref.json
