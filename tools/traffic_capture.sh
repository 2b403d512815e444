#!/bin/bash
# DRAM traffic + duration of every launch of ONE eager step of a bench config, on the current build:
#   bash tools/traffic_capture.sh [config] [outdir]    -> <outdir>/traffic_config<N>.csv  (copy it to profiles/r02/, tracked)
# bench.py parses the largest stage-T launch of this file into roofline.traffic (never a literal).
cfg=${1:-2}; out=${2:-gpurun_out}
mkdir -p "$out"
ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --csv --log-file "$out/traffic_config$cfg.csv" \
    python bench.py --config "$cfg" --steps 1 --warmup 3 --no-cuda-graph --no-cpu-baseline --no-train-probe --no-duplex-probe --no-fp32-convs \
    > "$out/traffic_config$cfg.log" 2>&1
echo "ncu rc=$? -> $out/traffic_config$cfg.csv ($(grep -c token_tc "$out/traffic_config$cfg.csv") token_tc rows)"
