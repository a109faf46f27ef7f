# round 2, GPU call M: small-batch FMA contraction (configs[0]), suite incl. the configs[4]-dimension parity test
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --timeout 1500 -rs ) > gpurun_out/m_gpu_tests.log 2>&1
tail -8 gpurun_out/m_gpu_tests.log
grep -E "configs\[4\] dims" gpurun_out/m_gpu_tests.log | cut -c1-400
for sb in 1 0; do
DADMM_SMALLB=$sb timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --workload cfg1 > gpurun_out/m_bench_cfg1_sb$sb.json 2> gpurun_out/m_bench_cfg1_sb$sb.err
DADMM_SMALLB=$sb timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --workload cfg1 --cuda-graph > gpurun_out/m_bench_cfg1_graph_sb$sb.json 2> gpurun_out/m_bench_cfg1_graph_sb$sb.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/m_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("m_bench_")[1], round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), j["gpu_launches"], j["loss_final"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict)})
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
