mkdir -p gpurun_out
python -m pytest tests/test_e2e_gpu.py tests/test_kernels_gpu.py -q -m gpu -x --tb=short 2>&1 | tail -4
python bench.py --no-cpu-baseline > gpurun_out/bench35.json 2> gpurun_out/bench35.err
cat gpurun_out/bench35.json | cut -c1-700
GANSPACE_B200_TIMELINE=1 python tools/phase_probe.py > gpurun_out/timeline35.log 2>&1
grep "== rep" gpurun_out/timeline35.log
