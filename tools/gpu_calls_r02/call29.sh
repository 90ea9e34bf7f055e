mkdir -p gpurun_out
python tools/ab_mapping.py tools/gpu_calls_r02/lib_4e907c2.so ganspace_b200/libganspace_b200.so > gpurun_out/ab29.log 2>&1
cat gpurun_out/ab29.log
python bench.py > gpurun_out/bench29.json 2> gpurun_out/bench29.err
cat gpurun_out/bench29.json
python -m pytest tests -q -m gpu -x --tb=short > gpurun_out/pytest29.log 2>&1
tail -15 gpurun_out/pytest29.log
