# round 2, GPU call D: GCN epilogue kernels, whole-step CUDA graphs, new defaults (fwd 4 CTAs/SM, 128-k first stage)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 -x ) > gpurun_out/d_gpu_tests.log 2>&1
tail -12 gpurun_out/d_gpu_tests.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/d_smoke.log 2>&1; tail -2 gpurun_out/d_smoke.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/d_bench_cfg4.json 2> gpurun_out/d_bench_cfg4.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg3 > gpurun_out/d_bench_cfg3.json 2> gpurun_out/d_bench_cfg3.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload cfg1 > gpurun_out/d_bench_cfg1.json 2> gpurun_out/d_bench_cfg1.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload cfg1 --cuda-graph > gpurun_out/d_bench_cfg1_graph.json 2> gpurun_out/d_bench_cfg1_graph.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg3 --cuda-graph > gpurun_out/d_bench_cfg3_graph.json 2> gpurun_out/d_bench_cfg3_graph.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/d_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("d_bench_")[1], round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), j["gpu_launches"], j["loss_final"], j["clocks"]["sm_mhz"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict)}, "fwd", round(r.get("step_fwd_hbm_frac",0),3), "bwd", round(r.get("step_bwd_hbm_frac",0),3), "step", round(r["step_hbm_frac"],3))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
timeout 300 python tools/bench_model3.py > gpurun_out/d_model3.txt 2>&1; tail -2 gpurun_out/d_model3.txt
timeout 300 python tools/bench_model3.py --cuda-graph > gpurun_out/d_model3_graph.txt 2>&1; tail -3 gpurun_out/d_model3_graph.txt
timeout 300 python tools/profile_model3.py > gpurun_out/d_model3_prof.txt 2>&1; tail -40 gpurun_out/d_model3_prof.txt
