# multi-GPU numbers after the stage kernels were split: torchrun N = 8, 4, 2 and the single-process C-ABI mode at 8 (verify metric only; the R1CS scaling is unchanged)
set -x
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 --skip-extra --no-cpu 2>gpurun_out/r2_b8.err | grep '^{' > gpurun_out/bench_r02_8gpu_verify.json; tail -2 gpurun_out/r2_b8.err; cut -c1-200 gpurun_out/bench_r02_8gpu_verify.json
python bench.py --abi-multi 8 --steps 10 --warmup 3 2>gpurun_out/r2_abi8.err | grep '^{' > gpurun_out/bench_r02_abi_multi_8gpu.json; tail -2 gpurun_out/r2_abi8.err; cut -c1-300 gpurun_out/bench_r02_abi_multi_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 5 --warmup 3 --skip-extra --no-cpu 2>gpurun_out/r2_b4.err | grep '^{' > gpurun_out/bench_r02_4gpu_verify.json; cut -c1-200 gpurun_out/bench_r02_4gpu_verify.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 5 --warmup 3 --skip-extra --no-cpu 2>gpurun_out/r2_b2.err | grep '^{' > gpurun_out/bench_r02_2gpu_verify.json; cut -c1-200 gpurun_out/bench_r02_2gpu_verify.json
python -m pytest tests -x -q -m gpu -k "multi_gpu" 2>&1 | tail -2
