set -x
nvidia-smi -L | wc -l
./tests/_hostemu/test_api 2>&1 | tail -2
python -m pytest tests -x -q -m gpu -k "multi_gpu" 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 2>gpurun_out/r2_b8.err | grep '^{' > gpurun_out/bench_r02_8gpu.json; tail -2 gpurun_out/r2_b8.err; cut -c1-200 gpurun_out/bench_r02_8gpu.json
python bench.py --abi-multi 8 --steps 5 --warmup 3 2>gpurun_out/r2_abi8.err | grep '^{' > gpurun_out/bench_r02_abi_multi_8gpu.json; tail -2 gpurun_out/r2_abi8.err; cut -c1-300 gpurun_out/bench_r02_abi_multi_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 5 --warmup 3 2>gpurun_out/r2_b4.err | grep '^{' > gpurun_out/bench_r02_4gpu.json; cut -c1-200 gpurun_out/bench_r02_4gpu.json
