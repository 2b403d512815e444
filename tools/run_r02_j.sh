mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_j.log; tail -25 gpurun_out/pytest_gpu_j.log
L=$PWD/gansformer-reproducibility-challenge_b200/libgf_attn_head.so
for rep in 1 2; do
echo "== simplex post new"; AB_POST=1 AB_MODES=default AB_ONLY=64,128,256 python tools/attn_bench.py 2>&1 | grep res=
echo "== simplex post prev"; GF_ATTN_LIB=$L AB_POST=1 AB_MODES=default AB_ONLY=64,128,256 python tools/attn_bench.py 2>&1 | grep res=
done
