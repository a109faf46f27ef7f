# round 2, GPU call F: half-length consensus gather A/B, model-#3 graphed step, inference sweep shard (configs[4])
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 ) > gpurun_out/f_gpu_tests.log 2>&1
tail -12 gpurun_out/f_gpu_tests.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; tail -2 gpurun_out/f_smoke.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_$name.json 2> gpurun_out/f_bench_$name.err; }
run fast DADMM_X=0
run exact DADMM_EXACT_ORDER=1
run fast_fwd5 DADMM_LEAN_MINB_FWD=5
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg3 > gpurun_out/f_bench_cfg3.json 2> gpurun_out/f_bench_cfg3.err
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload cfg5 --inference > gpurun_out/f_bench_cfg5_inference.json 2> gpurun_out/f_bench_cfg5_inference.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/f_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("f_bench_")[1], round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), j["gpu_launches"], j["loss_final"], j["clocks"]["sm_mhz"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict)}, "fwd", round(r.get("step_fwd_hbm_frac",0),3), "bwd", round(r.get("step_bwd_hbm_frac",0),3), "step", round(r["step_hbm_frac"],3))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
timeout 300 python tools/bench_model3.py > gpurun_out/f_model3.txt 2>&1; tail -2 gpurun_out/f_model3.txt
timeout 300 python tools/bench_model3.py --cuda-graph > gpurun_out/f_model3_graph.txt 2>&1; tail -3 gpurun_out/f_model3_graph.txt
