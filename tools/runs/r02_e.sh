# round 2, GPU call E: GCN fix, tensor-map TMA forward pipeline A/B, model-#3 step eager / graphed
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 ) > gpurun_out/e_gpu_tests.log 2>&1
tail -12 gpurun_out/e_gpu_tests.log
DADMM_FWD_PIPE=1 timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "chain or parity or baseline_shapes or modules" > gpurun_out/e_gpu_tests_pipe.log 2>&1
tail -4 gpurun_out/e_gpu_tests_pipe.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench_$name.json 2> gpurun_out/e_bench_$name.err; }
run default DADMM_X=0
run pipe_tmap16 DADMM_FWD_PIPE=1
run pipe_tmap8 DADMM_FWD_PIPE=1 DADMM_PIPE_WARPS=8
run pipe_rows DADMM_FWD_PIPE=1 DADMM_PIPE_TMAP=0
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload cfg1 --cuda-graph > gpurun_out/e_bench_cfg1_graph.json 2> gpurun_out/e_bench_cfg1_graph.err
DADMM_FWD_PIPE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg3 > gpurun_out/e_bench_cfg3_pipe.json 2> gpurun_out/e_bench_cfg3_pipe.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/e_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("e_bench_")[1], round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), j["gpu_launches"], j["loss_final"], j["clocks"]["sm_mhz"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict)}, "fwd", round(r.get("step_fwd_hbm_frac",0),3), "bwd", round(r.get("step_bwd_hbm_frac",0),3), "step", round(r["step_hbm_frac"],3))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
timeout 300 python tools/bench_model3.py > gpurun_out/e_model3.txt 2>&1; tail -2 gpurun_out/e_model3.txt
timeout 300 python tools/bench_model3.py --cuda-graph > gpurun_out/e_model3_graph.txt 2>&1; tail -3 gpurun_out/e_model3_graph.txt
timeout 300 python tools/profile_model3.py > gpurun_out/e_model3_prof.txt 2>&1; head -30 gpurun_out/e_model3_prof.txt | cut -c1-200
