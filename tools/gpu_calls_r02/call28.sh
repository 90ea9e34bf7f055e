mkdir -p gpurun_out
L="tools/gpu_calls_r02/lib_4e907c2.so tools/gpu_calls_r02/lib_candA.so ganspace_b200/libganspace_b200.so"
echo "== default env" > gpurun_out/ab28.log
python tools/ab_mapping.py $L >> gpurun_out/ab28.log 2>&1
echo "== CLUSTER=1" >> gpurun_out/ab28.log
GANSPACE_B200_MAPPING_CLUSTER=1 python tools/ab_mapping.py $L >> gpurun_out/ab28.log 2>&1
echo "== CLUSTER=4" >> gpurun_out/ab28.log
GANSPACE_B200_MAPPING_CLUSTER=4 python tools/ab_mapping.py $L >> gpurun_out/ab28.log 2>&1
python -m pytest tests/test_render_gpu.py tests/test_synthesis_gpu.py -q -m gpu --tb=long > gpurun_out/render28.log 2>&1
tail -5 gpurun_out/render28.log
cat gpurun_out/ab28.log
