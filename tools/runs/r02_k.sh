# round 2, GPU call K: prefetch A/B, interleaved twice (the boxes' clocks wander by +-5 % under the power cap)
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/k_bench_$name.json 2> gpurun_out/k_bench_$name.err; }
for rep in a b; do
run f0b0_$rep DADMM_LEVEL_PREFETCH_FWD=0 DADMM_LEVEL_PREFETCH_BWD=0
run f2b0_$rep DADMM_LEVEL_PREFETCH_FWD=2 DADMM_LEVEL_PREFETCH_BWD=0
run f2b1_$rep DADMM_LEVEL_PREFETCH_FWD=2 DADMM_LEVEL_PREFETCH_BWD=1
run f2b2_$rep DADMM_LEVEL_PREFETCH_FWD=2 DADMM_LEVEL_PREFETCH_BWD=2
run f1b1_$rep DADMM_LEVEL_PREFETCH_FWD=1 DADMM_LEVEL_PREFETCH_BWD=1
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/k_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("k_bench_")[1], round(j["value"]), round(j["ms_per_step"],3), j["clocks"]["sm_mhz"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict) and k in ("step_fwd","step_bwd","contract_tc","contract_stage1")}, "fwd", round(r.get("step_fwd_hbm_frac",0),3), "bwd", round(r.get("step_bwd_hbm_frac",0),3), "step", round(r["step_hbm_frac"],3))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
