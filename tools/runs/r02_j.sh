# round 2, GPU call J: software prefetch of the next row in the lean level kernels (A/B), suite
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/j_bench_$name.json 2> gpurun_out/j_bench_$name.err; }
run pf0 DADMM_LEVEL_PREFETCH=0
run pf1 DADMM_LEVEL_PREFETCH=1
run pf2 DADMM_LEVEL_PREFETCH=2
run pf0b DADMM_LEVEL_PREFETCH=0
run pf1_fwd4 DADMM_LEVEL_PREFETCH=1 DADMM_LEAN_MINB_FWD=4
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/j_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("j_bench_")[1], round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), j["loss_final"], j["clocks"]["sm_mhz"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict)}, "fwd", round(r.get("step_fwd_hbm_frac",0),3), "bwd", round(r.get("step_bwd_hbm_frac",0),3), "step", round(r["step_hbm_frac"],3))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
DADMM_LEVEL_PREFETCH=1 timeout 900 python -m pytest tests -m gpu -q --timeout 800 -k "chain or baseline_shapes or modules or graphs" > gpurun_out/j_tests_pf1.log 2>&1; tail -3 gpurun_out/j_tests_pf1.log
