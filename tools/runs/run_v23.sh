mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/r01_gpu_tests_v23.log 2>&1
tail -5 gpurun_out/r01_gpu_tests_v23.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_v23.log 2>&1; tail -2 gpurun_out/smoke_v23.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v23.json 2> gpurun_out/bench_v23.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pdl > gpurun_out/bench_v23_nopdl.json 2> gpurun_out/bench_v23_nopdl.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 512 > gpurun_out/bench_v23_b512.json 2> gpurun_out/bench_v23_b512.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 512 --no-pdl > gpurun_out/bench_v23_b512_nopdl.json 2> gpurun_out/bench_v23_b512_nopdl.err
python - <<'PY'
import json
for f in ("bench_v23","bench_v23_nopdl","bench_v23_b512","bench_v23_b512_nopdl"):
    try:
        j=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), j["gpu_launches"], j["loss_final"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict)})
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
