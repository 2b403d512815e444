mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$T --master-port 29711 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu_config2.out 2> gpurun_out/bench_8gpu_config2.err; echo "config2 rc=$?"
tail -n1 gpurun_out/bench_8gpu_config2.out | cut -c1-300
$T --master-port 29712 bench.py --gpus 8 --config 5 --steps 10 --warmup 3 --no-train-probe > gpurun_out/bench_8gpu_config5.out 2> gpurun_out/bench_8gpu_config5.err; echo "config5 rc=$?"
tail -n1 gpurun_out/bench_8gpu_config5.out | cut -c1-300
