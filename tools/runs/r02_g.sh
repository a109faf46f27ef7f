# round 2, GPU call G: final-candidate defaults, model-#3 with tensor-core linears, cfg5 forward level under ncu
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 ) > gpurun_out/g_gpu_tests.log 2>&1
tail -8 gpurun_out/g_gpu_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_cfg4.json 2> gpurun_out/g_bench_cfg4.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg3 > gpurun_out/g_bench_cfg3.json 2> gpurun_out/g_bench_cfg3.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/g_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("g_bench_")[1], round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), j["gpu_launches"], j["loss_final"], j["clocks"]["sm_mhz"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict)}, "fwd", round(r.get("step_fwd_hbm_frac",0),3), "bwd", round(r.get("step_bwd_hbm_frac",0),3), "step", round(r["step_hbm_frac"],3))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
timeout 300 python tools/bench_model3.py > gpurun_out/g_model3.txt 2>&1; tail -2 gpurun_out/g_model3.txt
timeout 300 python tools/bench_model3.py --cuda-graph > gpurun_out/g_model3_graph.txt 2>&1; tail -3 gpurun_out/g_model3_graph.txt
timeout 300 python tools/profile_model3.py > gpurun_out/g_model3_prof.txt 2>&1; head -26 gpurun_out/g_model3_prof.txt | cut -c1-200
timeout 600 python tools/profile_step.py --workload cfg5 --inference --K 3 > gpurun_out/g_profile_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'level_fwd' -s 4 -c 2 -f -o gpurun_out/r02_g_cfg5 python tools/profile_step.py --workload cfg5 --inference --K 3 > gpurun_out/g_ncu.log 2>&1
tail -3 gpurun_out/g_ncu.log; cat gpurun_out/g_profile_plain.log | tail -2
