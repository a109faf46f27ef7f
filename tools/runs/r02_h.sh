# round 2, GPU call H: P=100 forward level on 128-wide chunks (configs[4] inference shard), suite, model-#3 graphed step
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 ) > gpurun_out/h_gpu_tests.log 2>&1
tail -8 gpurun_out/h_gpu_tests.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/h_smoke.log 2>&1; tail -2 gpurun_out/h_smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload cfg5 --inference > gpurun_out/h_bench_cfg5_inference.json 2> gpurun_out/h_bench_cfg5_inference.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_cfg4.json 2> gpurun_out/h_bench_cfg4.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 512 > gpurun_out/h_bench_cfg4_b512.json 2> gpurun_out/h_bench_cfg4_b512.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/h_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("h_bench_")[1], round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), j["gpu_launches"], j["loss_final"], j["clocks"]["sm_mhz"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict)}, "fwd", round(r.get("step_fwd_hbm_frac",0),3), "bwd", round(r.get("step_bwd_hbm_frac",0),3), "step", round(r["step_hbm_frac"],3))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
timeout 300 python tools/bench_model3.py --cuda-graph > gpurun_out/h_model3_graph.txt 2>&1; tail -3 gpurun_out/h_model3_graph.txt
