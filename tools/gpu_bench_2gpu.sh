mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29721 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_config2.out 2> gpurun_out/bench_2gpu_config2.err; echo "config2 rc=$?"
tail -n1 gpurun_out/bench_2gpu_config2.out | cut -c1-400
$T --master-port 29722 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref.out 2> /dev/null; echo "ref rc=$?"; tail -n1 gpurun_out/bench_2gpu_ref.out | cut -c1-300
