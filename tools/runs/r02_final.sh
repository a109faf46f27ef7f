# round 2, final evidence run on one B200: suite, smoke, the bench line with its CPU legs, the reference arm, side configs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/final_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 -rs ) > gpurun_out/final_gpu_tests.log 2>&1
tail -8 gpurun_out/final_gpu_tests.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/final_bench_cfg4_reference_arm.json 2> gpurun_out/final_bench_cfg4_reference_arm.err
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench_cfg4.json 2> gpurun_out/final_bench_cfg4.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg3 > gpurun_out/final_bench_cfg3.json 2> gpurun_out/final_bench_cfg3.err
timeout 300 python bench.py --steps 20 --warmup 3 --workload cfg1 > gpurun_out/final_bench_cfg1.json 2> gpurun_out/final_bench_cfg1.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload cfg1 --cuda-graph > gpurun_out/final_bench_cfg1_graph.json 2> gpurun_out/final_bench_cfg1_graph.err
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload cfg5 --inference > gpurun_out/final_bench_cfg5_inference.json 2> gpurun_out/final_bench_cfg5_inference.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 512 > gpurun_out/final_bench_cfg4_b512.json 2> gpurun_out/final_bench_cfg4_b512.err
DADMM_EXACT_ORDER=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_cfg4_exact_order.json 2> gpurun_out/final_bench_cfg4_exact_order.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        if j.get("impl") == "reference":
            print(f.split("final_bench_")[1], "REFERENCE ARM", j["value"], j["ms_per_step"], j["cpu_baseline"]); continue
        r=j["roofline"]
        print(f.split("final_bench_")[1], round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), "fresh", j["e2e"].get("fresh_graphs") and round(j["e2e"]["fresh_graphs"]["ms_per_step"],2), j["gpu_launches"], j["loss_final"], j["clocks"]["sm_mhz"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict)}, "top", r["kernel"][:18], round(r["frac"],3), "fwd", round(r.get("step_fwd_hbm_frac",0),3), "bwd", round(r.get("step_bwd_hbm_frac",0),3), "step", round(r["step_hbm_frac"],3), "cpu", j.get("cpu_baseline",{}).get("value"), j.get("cpu_baseline",{}).get("kind"))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
timeout 300 python tools/bench_model3.py > gpurun_out/final_model3.txt 2>&1; tail -3 gpurun_out/final_model3.txt | cut -c1-400
timeout 300 python tools/bench_model3.py --cuda-graph > gpurun_out/final_model3_graph.txt 2>&1; tail -3 gpurun_out/final_model3_graph.txt | cut -c1-400
