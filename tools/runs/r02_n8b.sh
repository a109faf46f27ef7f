# round 2, second 8-GPU call: strong-scaling bench line of configs[3] on the final build (tile-width rule in)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/n8b_bench_cfg4.json 2> gpurun_out/n8b_bench_cfg4.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/n8b_bench_cfg4_n1.json 2> gpurun_out/n8b_bench_cfg4_n1.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/n8b_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("gpurun_out/")[1], j["n_gpus"], round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), j["loss_final"], j["clocks"]["sm_mhz"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict)}, "step", round(r["step_hbm_frac"],3))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
