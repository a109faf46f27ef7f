# round 2, GPU call AA: how much of a 512-problem step (the 8-GPU per-rank batch) is host bubbles?  eager vs one CUDA graph
mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/aa_bench_$name.json 2>> gpurun_out/aa_err.txt; }
run b512_eager_a --batch 512
run b512_graph_a --batch 512 --cuda-graph
run b512_eager_b --batch 512
run b512_graph_b --batch 512 --cuda-graph
run b1024_eager --batch 1024
run b1024_graph --batch 1024 --cuda-graph
DADMM_F16_NT_WAVES=3 run cfg3_graph_w3 --workload cfg3 --cuda-graph
run cfg3_graph_w6 --workload cfg3 --cuda-graph
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/aa_bench_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("aa_bench_")[1], round(j["value"]), round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["ms_per_step"], 4), j["clocks"]["sm_mhz"], round(j["kernel_breakdown_ms"]["step_total_ms"], 3))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -3 gpurun_out/aa_err.txt
