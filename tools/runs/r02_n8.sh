# round 2, 8-GPU call: strong-scaling bench line (configs[3]) and the inference sweep of configs[4] (16384 problems over 8 GPUs)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/n8_bench_cfg4.json 2> gpurun_out/n8_bench_cfg4.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 3 --workload cfg5 --inference --batch 16384 > gpurun_out/n8_bench_cfg5_inference.json 2> gpurun_out/n8_bench_cfg5_inference.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/n4_bench_cfg4.json 2> gpurun_out/n4_bench_cfg4.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/n[48]_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("gpurun_out/")[1], j["n_gpus"], round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), j["loss_final"], j["clocks"]["sm_mhz"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict)}, "step", round(r["step_hbm_frac"],3))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
