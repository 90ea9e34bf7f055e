mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 180 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-parity 2>gpurun_out/err32.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['sections_ms_per_step'])" || tail -3 gpurun_out/err32.log; }
run A=1
run GANSPACE_B200_CHAIN_PERSISTENT=1
run GANSPACE_B200_RNG_GROUPS=1,3,7 GANSPACE_B200_STATS_FIRST=1
run GANSPACE_B200_RNG_GROUPS=2,7 GANSPACE_B200_STATS_FIRST=2
run GANSPACE_B200_RNG_GROUPS=1,3,7 GANSPACE_B200_STATS_FIRST=1 GANSPACE_B200_CHAIN_PERSISTENT=1
run GANSPACE_B200_STATS_BLOCK=5
run GANSPACE_B200_STATS_BLOCK=7 GANSPACE_B200_RNG_GROUPS=1,3,7 GANSPACE_B200_STATS_FIRST=1
GANSPACE_B200_TIMELINE=1 python tools/phase_probe.py > gpurun_out/timeline32.log 2>&1
