mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/ncu_bench33.log 2>&1
tail -2 gpurun_out/ncu_bench33.log | cut -c1-300
wc -l gpurun_out/launches_r02.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mapping_layer_tc_kernel -s 20 -c 1 -f -o gpurun_out/ncu_mapping_r02 python tools/bench_mapping.py > gpurun_out/ncu_map33.log 2>&1
tail -3 gpurun_out/ncu_map33.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
