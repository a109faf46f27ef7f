# round 2, GPU call I: suite on the candidate-final build, ncu full-set + launch list, backward-level occupancy A/B
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 ) > gpurun_out/i_gpu_tests.log 2>&1
tail -6 gpurun_out/i_gpu_tests.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench_$name.json 2> gpurun_out/i_bench_$name.err; }
run default DADMM_X=0
run bwd8w3 DADMM_LEVEL_WARPS=8 DADMM_LEAN_MINB_BWD=3
run bwd8w4 DADMM_LEVEL_WARPS=8 DADMM_LEAN_MINB_BWD=4
run bwd10w2 DADMM_LEAN_MINB_BWD=2
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/i_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("i_bench_")[1], round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), "fresh", j["e2e"].get("fresh_graphs") and round(j["e2e"]["fresh_graphs"]["ms_per_step"],2), j["loss_final"], j["clocks"]["sm_mhz"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict)}, "fwd", round(r.get("step_fwd_hbm_frac",0),3), "bwd", round(r.get("step_bwd_hbm_frac",0),3), "step", round(r["step_hbm_frac"],3))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
timeout 300 python tools/profile_step.py --K 3 > gpurun_out/i_profile_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'level_fwd|level_bwd|contract_f16' -s 16 -c 16 -f -o gpurun_out/r02_i_full python tools/profile_step.py --K 3 > gpurun_out/i_ncu.log 2>&1
tail -2 gpurun_out/i_ncu.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench_plain_for_ncu.json 2>/dev/null &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_i_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/i_ncu_launches.log 2>&1
tail -2 gpurun_out/i_ncu_launches.log; wc -l gpurun_out/r02_i_launches.csv
