# round 2, GPU call Y: skinny-batch contraction kernel and small-batch level tiles: suite, configs[0] eager / graph, A/B against the old kernels
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x ) > gpurun_out/y_gpu_tests.log 2>&1
tail -4 gpurun_out/y_gpu_tests.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/y_bench_cfg1_$name.json 2>> gpurun_out/y_err.txt; }
EXTRA="" run eager X=1
EXTRA="--cuda-graph" run graph X=1
EXTRA="--cuda-graph" run graph_oldgemm DADMM_SKINNY=0
EXTRA="--cuda-graph" run graph_oldtiles DADMM_STEP_TB=4
EXTRA="--cuda-graph" run graph_b X=1
timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 3 > gpurun_out/y_bench_cfg1_full.json 2>> gpurun_out/y_err.txt
timeout 300 python tools/bench_model3.py --cuda-graph > gpurun_out/y_model3_graph.txt 2>&1; tail -1 gpurun_out/y_model3_graph.txt | cut -c1-200
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/y_bench_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("y_bench_")[1], round(j["value"]), round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["ms_per_step"], 4), j["loss_final"], {k: (v["ms"], v["launches"]) for k, v in j["kernel_breakdown_ms"].items() if isinstance(v, dict)})
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -5 gpurun_out/y_err.txt
