# round 2, GPU call AF: config 4 with two problems per FORWARD level tile (the backward keeps one: its two tiles do not fit twice)
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/af_bench_cfg4_$name.json 2>> gpurun_out/af_err.txt; }
for rep in a b; do
run tb1_$rep X=1
run tb2_$rep DADMM_STEP_TB=2
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/af_bench_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("af_bench_")[1], round(j["value"]), round(j["ms_per_step"], 3), j["clocks"]["sm_mhz"], j["loss_final"], {k: v["ms"] for k, v in j["kernel_breakdown_ms"].items() if isinstance(v, dict)})
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -3 gpurun_out/af_err.txt
