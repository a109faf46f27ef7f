# round 2, GPU call AD: configs[2] (P=20: 20-row tiles on 8 warps) with two / four problems per level tile
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --cuda-graph > gpurun_out/ad_bench_cfg3_$name.json 2>> gpurun_out/ad_err.txt; }
for rep in a b; do
run tb1_$rep X=1
run tb2_$rep DADMM_STEP_TB=2
run tb4_$rep DADMM_STEP_TB=4
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ad_bench_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("ad_bench_")[1], round(j["value"]), round(j["ms_per_step"], 4), j["loss_final"], {k: v["ms"] for k, v in j["kernel_breakdown_ms"].items() if isinstance(v, dict)})
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -3 gpurun_out/ad_err.txt
