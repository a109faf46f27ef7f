# round 2, GPU call O: L2 prefetch of the next tile's seeds in the contraction epilogue (A/B, interleaved)
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/o_bench_$name.json 2> gpurun_out/o_bench_$name.err; }
for rep in a b c; do
run sp1_$rep DADMM_SEED_PREFETCH=1
run sp0_$rep DADMM_SEED_PREFETCH=0
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/o_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("o_bench_")[1], round(j["value"]), round(j["ms_per_step"],3), j["clocks"]["sm_mhz"], j["loss_final"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict) and k in ("step_fwd","step_bwd","contract_tc","contract_stage1")})
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
timeout 600 python -m pytest tests -m gpu -q --timeout 500 -k "parity or chain or baseline_shapes" > gpurun_out/o_tests.log 2>&1; tail -3 gpurun_out/o_tests.log
