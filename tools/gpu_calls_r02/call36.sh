mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -x --tb=short 2>&1 | tail -4 | tee gpurun_out/pytest36_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench36_2gpu.json 2> gpurun_out/bench36_2gpu.err
cat gpurun_out/bench36_2gpu.json | cut -c1-600; tail -2 gpurun_out/bench36_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 1 --warmup 1 --config 5 --no-cpu-baseline > gpurun_out/bench36_2gpu_c5.json 2> gpurun_out/bench36_2gpu_c5.err
cat gpurun_out/bench36_2gpu_c5.json | cut -c1-400; tail -2 gpurun_out/bench36_2gpu_c5.err
