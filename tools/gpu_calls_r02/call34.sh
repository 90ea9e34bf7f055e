mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "split or normal" --tb=short 2>&1 | tail -3
run() { echo "== $*"; env "$@" timeout 180 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-parity 2>gpurun_out/err34.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['sections_ms_per_step'])" || tail -3 gpurun_out/err34.log; }
run A=1
run GANSPACE_B200_RNG_FIRST_PARTS=8
run GANSPACE_B200_RNG_GROUPS=3,7 GANSPACE_B200_STATS_FIRST=3
run GANSPACE_B200_RNG_GROUPS=2,7 GANSPACE_B200_STATS_FIRST=2
run A=2
python -m pytest tests/test_e2e_gpu.py -q -m gpu -x --tb=short 2>&1 | tail -3
GANSPACE_B200_TIMELINE=1 python tools/phase_probe.py 2>&1 | grep -v "chain done" | grep "ms  \|== rep" | head -30
bash tools/gpu_calls_r02/call33.sh
