# round 2, GPU call A: full GPU suite on the inherited build, baselines of every workload, first ncu of the round
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/a_smi.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 900 ) > gpurun_out/a_gpu_tests.log 2>&1
tail -5 gpurun_out/a_gpu_tests.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/a_smoke.log 2>&1; tail -2 gpurun_out/a_smoke.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_cfg4.json 2> gpurun_out/a_bench_cfg4.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg3 > gpurun_out/a_bench_cfg3.json 2> gpurun_out/a_bench_cfg3.err
for tb in 2 4; do
DADMM_STEP_TB=$tb timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg3 > gpurun_out/a_bench_cfg3_tb$tb.json 2> gpurun_out/a_bench_cfg3_tb$tb.err
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload cfg1 > gpurun_out/a_bench_cfg1.json 2> gpurun_out/a_bench_cfg1.err
timeout 300 python tools/bench_model3.py > gpurun_out/a_model3.txt 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/a_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), j["gpu_launches"], j["loss_final"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict)})
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
cat gpurun_out/a_model3.txt | tail -3
# ncu: full set on the second training step of a K=3 run (first step = 16 matching launches skipped)
timeout 300 python tools/profile_step.py --K 3 > gpurun_out/a_profile_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'level_fwd_kernel|level_bwd_kernel|contract_f16' -s 16 -c 16 -f -o gpurun_out/r02_a_full python tools/profile_step.py --K 3 > gpurun_out/a_ncu.log 2>&1
tail -3 gpurun_out/a_ncu.log
ls -la gpurun_out | tail -20
