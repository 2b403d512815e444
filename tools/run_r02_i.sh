mkdir -p gpurun_out
L=$PWD/gansformer-reproducibility-challenge_b200/libgf_attn_head.so
for rep in 1 2; do
echo "== new (working tree)"; DL_ONLY=128,256 python tools/duplex_layers.py 2>&1 | grep res=
echo "== prev (HEAD)"; GF_ATTN_LIB=$L DL_ONLY=128,256 python tools/duplex_layers.py 2>&1 | grep res=
done
echo "== simplex post new"; AB_POST=1 AB_MODES=default AB_ONLY=64,128,256 python tools/attn_bench.py 2>&1 | grep res=
echo "== simplex post prev"; GF_ATTN_LIB=$L AB_POST=1 AB_MODES=default AB_ONLY=64,128,256 python tools/attn_bench.py 2>&1 | grep res=
echo "== simplex post new"; AB_POST=1 AB_MODES=default AB_ONLY=64,128,256 python tools/attn_bench.py 2>&1 | grep res=
echo "== simplex post prev"; GF_ATTN_LIB=$L AB_POST=1 AB_MODES=default AB_ONLY=64,128,256 python tools/attn_bench.py 2>&1 | grep res=
