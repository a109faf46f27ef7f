# round 2, GPU call AE: two problems per level tile for 16..31 agents (configs[2]) with the label-free loss sums kept: suite, configs[2] eager / graph
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/ae_gpu_tests.log 2>&1
tail -4 gpurun_out/ae_gpu_tests.log
run() { name=$1; shift; timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ae_bench_cfg3_$name.json 2>> gpurun_out/ae_err.txt; }
run eager
run graph --cuda-graph
DADMM_STEP_TB=1 run graph_tb1 --cuda-graph
run graph_b --cuda-graph
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ae_bench_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("ae_bench_")[1], round(j["value"]), round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["ms_per_step"], 4), j["loss_final"], {k: v["ms"] for k, v in j["kernel_breakdown_ms"].items() if isinstance(v, dict)})
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -3 gpurun_out/ae_err.txt
