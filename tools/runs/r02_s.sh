# round 2, GPU call S: wave threshold of the contraction's batch-tile width at the per-rank batches of 8 / 4 / 2 GPUs
mkdir -p gpurun_out
run() { name=$1; b=$2; shift; shift; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --batch $b > gpurun_out/s_bench_$name.json 2> gpurun_out/s_err.txt; }
for rep in a b; do
for b in 512 1024 2048; do
run w3_${b}_$rep $b DADMM_F16_NT_WAVES=3
run w6_${b}_$rep $b DADMM_F16_NT_WAVES=6
run w12_${b}_$rep $b DADMM_F16_NT_WAVES=12
done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/s_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("s_bench_")[1], round(j["value"]), round(j["ms_per_step"],3), j["clocks"]["sm_mhz"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict) and k in ("contract_tc","contract_stage1")})
    except Exception as e:
        print(f, "FAILED", e)
PY
