# round 2, GPU call AG: ten warps per CTA in the lean forward level (rows split evenly at P = 50 / 20) -- config 4 and configs[2]
mkdir -p gpurun_out
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ag_bench_$name.json 2>> gpurun_out/ag_err.txt; }
for rep in a b; do
run cfg4_w8_$rep cfg4 X=1
run cfg4_w10_$rep cfg4 DADMM_FWD_WARPS=10
done
run cfg3_w8 cfg3 X=1
run cfg3_w10 cfg3 DADMM_FWD_WARPS=10
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ag_bench_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("ag_bench_")[1], round(j["value"]), round(j["ms_per_step"], 3), j["clocks"]["sm_mhz"], j["loss_final"], {k: v["ms"] for k, v in j["kernel_breakdown_ms"].items() if isinstance(v, dict) and k.startswith("step")})
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -3 gpurun_out/ag_err.txt
