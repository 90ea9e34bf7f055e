mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench37_4gpu.json 2> gpurun_out/bench37_4gpu.err
cat gpurun_out/bench37_4gpu.json | cut -c1-300; tail -2 gpurun_out/bench37_4gpu.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 4 --steps 1 --warmup 1 --config 5 --no-cpu-baseline > gpurun_out/bench37_4gpu_c5.json 2> gpurun_out/bench37_4gpu_c5.err
cat gpurun_out/bench37_4gpu_c5.json | cut -c1-300; tail -2 gpurun_out/bench37_4gpu_c5.err
