# round 2, GPU call Q: lean streaming kernel for backward level 0 -- suite + bench
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --timeout 1500 ) > gpurun_out/q_gpu_tests.log 2>&1
tail -5 gpurun_out/q_gpu_tests.log
for rep in a b; do
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/q_bench_new_$rep.json 2> gpurun_out/q_err.txt
DADMM_BWD_GEN=1 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/q_bench_bwdgen1_$rep.json 2> gpurun_out/q_err.txt
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 512 > gpurun_out/q_bench_b512.json 2> gpurun_out/q_err.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/q_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("q_bench_")[1], round(j["value"]), round(j["ms_per_step"],3), j["clocks"]["sm_mhz"], j["loss_final"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict) and k in ("step_fwd","step_bwd","contract_tc","contract_stage1")})
    except Exception as e:
        print(f, "FAILED", e)
PY
