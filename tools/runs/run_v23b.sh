mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/bench_default_v23.json 2> gpurun_out/bench_default_v23.err
timeout 120 python bench.py --workload cfg1 --no-cpu-baseline --steps 30 > gpurun_out/bench_cfg1_v23.json 2> gpurun_out/bench_cfg1_v23.err
timeout 120 python bench.py --workload cfg3 --no-cpu-baseline --steps 10 > gpurun_out/bench_cfg3_v23.json 2> gpurun_out/bench_cfg3_v23.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r01_v23.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_v23.log 2>&1
python - <<'PY'
import json
for f in ("bench_default_v23","bench_cfg1_v23","bench_cfg3_v23"):
    try:
        j=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), j["gpu_launches"], j["clocks"], j.get("cpu_baseline"), j["roofline"]["frac"])
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
tail -3 gpurun_out/ncu_launches_v23.log; wc -l gpurun_out/launches_r01_v23.csv
