# round 2, GPU call N: small-batch kernel (unrolled) A/B at configs[0]; split-K weight gradients in model #3
mkdir -p gpurun_out
for sb in 1 0 1 0; do
DADMM_SMALLB=$sb timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --workload cfg1 --cuda-graph > gpurun_out/n_bench_cfg1_graph_sb${sb}_$RANDOM.json 2> gpurun_out/n_err.txt
done
DADMM_SMALLB=1 timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --workload cfg1 > gpurun_out/n_bench_cfg1_sb1.json 2> gpurun_out/n_err.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/n_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("n_bench_")[1], round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), j["loss_final"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict)})
    except Exception as e:
        print(f, "FAILED", e)
PY
timeout 600 python -m pytest tests -m gpu -q --timeout 500 -k "gcn or model3 or parity or chain" > gpurun_out/n_tests.log 2>&1; tail -3 gpurun_out/n_tests.log
timeout 300 python tools/bench_model3.py --cuda-graph > gpurun_out/n_model3_graph.txt 2>&1; tail -3 gpurun_out/n_model3_graph.txt | cut -c1-300
timeout 300 python tools/bench_model3.py > gpurun_out/n_model3.txt 2>&1; tail -3 gpurun_out/n_model3.txt | cut -c1-300
