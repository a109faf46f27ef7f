# round 2, GPU call AC: ABI 7 (kept operator copy for per-iteration contractions), skinny-kernel shape sweep: full suite, smoke, model #3, default bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/ac_gpu_tests.log 2>&1
tail -4 gpurun_out/ac_gpu_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/ac_smoke.log 2>&1; tail -1 gpurun_out/ac_smoke.log | cut -c1-300
timeout 300 python tools/bench_model3.py --cuda-graph > gpurun_out/ac_model3_graph.txt 2>&1; tail -1 gpurun_out/ac_model3_graph.txt | cut -c1-250
DADMM_OP_SPLIT_CACHE=0 timeout 300 python tools/bench_model3.py --cuda-graph > gpurun_out/ac_model3_graph_nocache.txt 2>&1; tail -1 gpurun_out/ac_model3_graph_nocache.txt | cut -c1-250
timeout 300 python tools/bench_model3.py > gpurun_out/ac_model3_eager.txt 2>&1; tail -1 gpurun_out/ac_model3_eager.txt | cut -c1-250
timeout 600 python bench.py > gpurun_out/ac_bench_cfg4.json 2> gpurun_out/ac_err.txt
python - <<'PY'
import json
j = json.loads(open("gpurun_out/ac_bench_cfg4.json").read().strip().splitlines()[-1])
print("cfg4", round(j["value"]), round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["ms_per_step"], 3), j["clocks"], j["cpu_baseline"]["value"])
PY
