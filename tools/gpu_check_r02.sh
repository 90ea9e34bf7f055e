mkdir -p gpurun_out
echo "== A/B mapping"; timeout 300 python tools/ab_mapping.py tools/_ab/lib_4e907c2.so tools/_ab/lib_146bfb4.so ganspace_b200/libganspace_b200.so 2>&1 | tail -8
echo "== render tests"; timeout 900 python -m pytest tests/test_render_gpu.py -q -m gpu 2>&1 | grep -v "^Reusing\|^Using\|^Feature\|^B=" | tail -40
