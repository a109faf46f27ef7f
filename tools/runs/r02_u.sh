# round 2, GPU call U: full -m gpu suite (with the guard-band tests), smoke, default bench, 512-problem bench on the final build
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/u_gpu_tests.log 2>&1
tail -4 gpurun_out/u_gpu_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/u_smoke.log 2>&1; tail -2 gpurun_out/u_smoke.log
timeout 600 python bench.py > gpurun_out/u_bench_cfg4.json 2> gpurun_out/u_err.txt; tail -c 600 gpurun_out/u_bench_cfg4.json
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --batch 512 > gpurun_out/u_bench_b512.json 2>> gpurun_out/u_err.txt
python - <<'PY'
import json
for f in ("u_bench_cfg4", "u_bench_b512"):
    j = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(j["value"]), round(j["ms_per_step"], 3), j["clocks"], j["e2e"]["value"])
PY
