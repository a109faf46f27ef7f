# round 2, GPU call R: batch-tile width of the contraction at the 8-GPU per-rank batch (512 problems), interleaved
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 512 > gpurun_out/r_bench_$name.json 2> gpurun_out/r_err.txt; }
for rep in a b; do
run auto_$rep DADMM_X=0
run nt128_$rep DADMM_F16_NT=128
run nt256_$rep DADMM_F16_NT=256
run nopdl_$rep DADMM_PDL=0
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r_bench_")[1], round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), j["clocks"]["sm_mhz"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict) and k in ("step_fwd","step_bwd","contract_tc","contract_stage1","loss","split")}, "sum", round(sum(v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict)),2))
    except Exception as e:
        print(f, "FAILED", e)
PY
