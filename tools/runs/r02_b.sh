# round 2, GPU call B: second-generation lean level kernels (packed fp32) -- parity suite, A/B timings, ncu
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 -x -s -k "baseline_shapes or guard_branch" ) > gpurun_out/b_new_tests.log 2>&1
tail -25 gpurun_out/b_new_tests.log
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_gpu_baseline_shapes.py --deselect tests/test_gpu_guard_branches.py ) > gpurun_out/b_gpu_tests.log 2>&1
tail -8 gpurun_out/b_gpu_tests.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/b_smoke.log 2>&1; tail -2 gpurun_out/b_smoke.log
run() { # name, env..., -- args
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_$name.json 2> gpurun_out/b_bench_$name.err
}
run default DADMM_X=0
run gen1 DADMM_LEVEL_GEN=1
run fwd4 DADMM_LEAN_MINB_FWD=4
run fwd3 DADMM_LEAN_MINB_FWD=3
run bwd2 DADMM_LEAN_MINB_BWD=2
run bwd8w3 DADMM_LEVEL_WARPS=8 DADMM_LEAN_MINB_BWD=3
run bwd8w4 DADMM_LEVEL_WARPS=8 DADMM_LEAN_MINB_BWD=4
run kbc2 DADMM_F16_KBC=2
run kbc4 DADMM_F16_KBC=4
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg3 > gpurun_out/b_bench_cfg3.json 2> gpurun_out/b_bench_cfg3.err
DADMM_STEP_TB=2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg3 > gpurun_out/b_bench_cfg3_tb2.json 2> gpurun_out/b_bench_cfg3_tb2.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/b_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("b_bench_")[1], round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), "fresh", j["e2e"].get("fresh_graphs") and round(j["e2e"]["fresh_graphs"]["ms_per_step"],2), j["loss_final"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict)}, "top", r["kernel"], round(r["frac"],3), "fwd", round(r.get("step_fwd_hbm_frac",0),3), "bwd", round(r.get("step_bwd_hbm_frac",0),3), "step", round(r["step_hbm_frac"],3))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
timeout 300 python tools/profile_step.py --K 3 > gpurun_out/b_profile_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'level_fwd_lean|level_bwd_lean' -s 4 -c 4 -f -o gpurun_out/r02_b_full python tools/profile_step.py --K 3 > gpurun_out/b_ncu.log 2>&1
tail -3 gpurun_out/b_ncu.log
