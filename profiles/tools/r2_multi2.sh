set -x
nvidia-smi -L
python -m pytest tests -x -q -m gpu -k "multi_gpu or cpp_api" 2>&1 | tail -4
./tests/_hostemu/test_api 2>&1 | tail -3
python bench.py --abi-multi 2 --steps 3 --warmup 3 2>gpurun_out/r2_abi2.err | tee gpurun_out/r2_abi_multi_2gpu.json | cut -c1-400; tail -3 gpurun_out/r2_abi2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2>gpurun_out/r2_b2.err | tee gpurun_out/r2_bench_2gpu.json | cut -c1-600; tail -3 gpurun_out/r2_b2.err
