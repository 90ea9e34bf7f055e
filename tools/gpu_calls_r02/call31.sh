mkdir -p gpurun_out
python tools/rng_probe.py 101 10000 8 > gpurun_out/rng31.log 2>&1
python tools/rng_probe.py 101 10000 1 >> gpurun_out/rng31.log 2>&1
cat gpurun_out/rng31.log
python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "rng or normal or mt or split or sample" --tb=short 2>&1 | tail -3
run() { echo "== groups=$1 free=$2 streams=$3"; GANSPACE_B200_RNG_GROUPS=$1 GANSPACE_B200_LAZY_FREE_SMS=$2 GANSPACE_B200_RNG_STREAMS=$3 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['sections_ms_per_step'])"; }
run 4,7 72 1
run 4,6 64 1
run 3,6 64 1
run 4,8 80 1
run 2,7 72 1
run 4,5 56 1
