# round 2, GPU call C: pipelined forward level kernel -- parity suite, A/B timings, ncu
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 -x ) > gpurun_out/c_gpu_tests.log 2>&1
tail -8 gpurun_out/c_gpu_tests.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/c_smoke.log 2>&1; tail -2 gpurun_out/c_smoke.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_$name.json 2> gpurun_out/c_bench_$name.err; }
run pipe16 DADMM_X=0
run pipe12 DADMM_PIPE_WARPS=12
run pipe8 DADMM_PIPE_WARPS=8
run nopipe_fwd4 DADMM_FWD_PIPE=0 DADMM_LEAN_MINB_FWD=4
run pipe16_bwdgen1 DADMM_BWD_GEN=1
run pipe16_kbc4 DADMM_F16_KBC=4
runw() { name=$1; wl=$2; shift; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/c_bench_$name.json 2> gpurun_out/c_bench_$name.err; }
runw cfg3_pipe cfg3 DADMM_X=0
runw cfg3_pipe_tb1 cfg3 DADMM_PIPE_TB=1
runw cfg3_pipe_tb4 cfg3 DADMM_PIPE_TB=4
runw cfg3_nopipe cfg3 DADMM_FWD_PIPE=0
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("c_bench_")[1], round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), j["loss_final"], j["clocks"]["sm_mhz"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict)}, "fwd", round(r.get("step_fwd_hbm_frac",0),3), "bwd", round(r.get("step_bwd_hbm_frac",0),3), "step", round(r["step_hbm_frac"],3))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
timeout 300 python tools/profile_step.py --K 3 > gpurun_out/c_profile_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'level_fwd_pipe' -s 2 -c 2 -f -o gpurun_out/r02_c_full python tools/profile_step.py --K 3 > gpurun_out/c_ncu.log 2>&1
tail -3 gpurun_out/c_ncu.log
