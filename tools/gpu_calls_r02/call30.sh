mkdir -p gpurun_out
GANSPACE_B200_TIMING=1 timeout 600 python bench.py --config 5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench30_c5.json 2> gpurun_out/bench30_c5.err
cat gpurun_out/bench30_c5.json; tail -3 gpurun_out/bench30_c5.err
GANSPACE_B200_BIGD_CHAIN=lanczos timeout 600 python bench.py --config 5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench30_c5_lanczos.json 2> gpurun_out/bench30_c5_lanczos.err
cat gpurun_out/bench30_c5_lanczos.json; tail -3 gpurun_out/bench30_c5_lanczos.err
timeout 300 python bench.py --config 3 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench30_c3.json 2> gpurun_out/bench30_c3.err
cat gpurun_out/bench30_c3.json
timeout 300 python bench.py --config 4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench30_c4.json 2> gpurun_out/bench30_c4.err
cat gpurun_out/bench30_c4.json
timeout 300 python bench.py --config 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench30_c1.json 2> gpurun_out/bench30_c1.err
cat gpurun_out/bench30_c1.json
