# the driver's own command at 8 GPUs (verify + r1cs + secondary in one line) after the stage split and the long-row kernel
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 2>gpurun_out/r2_b8.err | grep '^{' > gpurun_out/bench_r02_8gpu.json; tail -2 gpurun_out/r2_b8.err; cut -c1-200 gpurun_out/bench_r02_8gpu.json
