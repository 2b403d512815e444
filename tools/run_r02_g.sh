set -x
mkdir -p gpurun_out
SAN_ONLY=centroid_tc/C128 timeout 300 compute-sanitizer --tool synccheck --print-limit 3 python tools/sanitize_driver.py > gpurun_out/synccheck_iso2.log 2>&1; grep -v "Host Frame" gpurun_out/synccheck_iso2.log | head -12
N="ncu --profile-from-start off --set full --clock-control none --import-source on"
NL_POST=2 $N -k regex:token_tc -o gpurun_out/ncu_full_token_tc_r02_res256_C128_K16_postop_rgb python tools/ncu_layer.py > /dev/null 2>&1
NL_RES=128 NL_C=256 NL_K=32 NL_B=64 NL_DUPLEX=1 NL_POST=1 $N -k regex:centroid_tc -o gpurun_out/ncu_full_centroid_tc_r02_res128_C256_K32 python tools/ncu_layer.py > /dev/null 2>&1
NL_RES=256 NL_C=128 NL_K=32 NL_B=64 NL_DUPLEX=1 NL_POST=1 $N -k regex:centroid_tc -o gpurun_out/ncu_full_centroid_tc_r02_res256_C128_K32 python tools/ncu_layer.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
