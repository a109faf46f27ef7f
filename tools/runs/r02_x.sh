# round 2, GPU call X: ncu launch list of the configs[0] training step (where do its 1.2 ms go?)
mkdir -p gpurun_out
timeout 300 python bench.py --workload cfg1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/x_bench_cfg1.json 2> gpurun_out/x_err.txt &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x_cfg1_launches.csv python bench.py --workload cfg1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/x_ncu.log 2>&1
tail -2 gpurun_out/x_ncu.log | cut -c1-300; wc -l gpurun_out/x_cfg1_launches.csv
