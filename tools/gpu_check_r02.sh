mkdir -p gpurun_out
echo "== chain default"; timeout 200 python tools/chain_probe.py 30 2>&1 | tail -4
echo "== chain dbl=10"; GANSPACE_B200_SUBSPACE_DBL=10 timeout 200 python tools/chain_probe.py 30 2>&1 | tail -4
echo "== chain blocked"; GANSPACE_B200_SUBSPACE_CHOL=blocked timeout 200 python tools/chain_probe.py 30 2>&1 | tail -4
echo "== chain dbl=10 blocked"; GANSPACE_B200_SUBSPACE_DBL=10 GANSPACE_B200_SUBSPACE_CHOL=blocked timeout 200 python tools/chain_probe.py 30 2>&1 | tail -4
echo "== mapping pair"; GANSPACE_B200_MAPPING_PAIR=1 timeout 60 python tools/bench_mapping.py 2>&1 | tail -1 | cut -c1-330
echo "== mapping nopair"; timeout 60 python tools/bench_mapping.py 2>&1 | tail -1 | cut -c1-330
echo "== render smoke"; timeout 300 python tools/render_smoke.py 2>&1 | tail -12
echo "== tests"; timeout 900 python -m pytest tests/test_render_gpu.py tests/test_synthesis_gpu.py tests/test_kernels_gpu.py tests/test_e2e_gpu.py -q -m gpu 2>&1 | tail -12
echo "== bench"; timeout 200 python bench.py --no-cpu-baseline 2>gpurun_out/bench6.err | tee gpurun_out/bench6.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['sections_ms_per_step'], d['parity']['ok'], d['roofline']['isolated'])"; tail -3 gpurun_out/bench6.err
