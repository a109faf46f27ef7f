# round 2, GPU call L: chunks per CTA of the lean level kernels (A/B, interleaved)
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/l_bench_$name.json 2> gpurun_out/l_bench_$name.err; }
for rep in a b; do
run base_$rep DADMM_X=0
run bcs2_$rep DADMM_BWD_CSPLIT=2
run bcs4_$rep DADMM_BWD_CSPLIT=4
run bcs8_$rep DADMM_BWD_CSPLIT=8
run fcpc1_$rep DADMM_FWD_CHUNKS_PER_CTA=1
run fcpc4_$rep DADMM_FWD_CHUNKS_PER_CTA=4
run fcpc8_$rep DADMM_FWD_CHUNKS_PER_CTA=8
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/l_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("l_bench_")[1], round(j["value"]), round(j["ms_per_step"],3), j["clocks"]["sm_mhz"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict) and k in ("step_fwd","step_bwd","contract_tc","contract_stage1")}, "fwd", round(r.get("step_fwd_hbm_frac",0),3), "bwd", round(r.get("step_bwd_hbm_frac",0),3))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
