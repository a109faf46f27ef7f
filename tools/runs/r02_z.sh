# round 2, GPU call Z: configs[2] under the tile-width threshold variants (its single-stage contraction has n_in = 256, 4.3 waves of 256-wide tiles)
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/z_bench_cfg3_$name.json 2>> gpurun_out/z_err.txt; }
for rep in a b; do
run w3_$rep DADMM_F16_NT_WAVES=3
run w6_$rep X=1
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/z_bench_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("z_bench_")[1], round(j["value"]), round(j["ms_per_step"], 4), "e2e", round(j["e2e"]["ms_per_step"], 4), j["clocks"]["sm_mhz"], {k: (v["ms"], v["launches"]) for k, v in j["kernel_breakdown_ms"].items() if isinstance(v, dict)})
    except Exception as e:
        print(f, "FAILED", e)
PY
