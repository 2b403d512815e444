set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_f_2gpu.json 2> gpurun_out/bench_f_2gpu.err; echo "rc=$?"; tail -c 600 gpurun_out/bench_f_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_f_2gpu_ref.json 2> gpurun_out/bench_f_2gpu_ref.err; echo "rc=$?"
