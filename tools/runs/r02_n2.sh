# round 2, 2-GPU call: same-noise rank parity, strong-scaling bench line, inference sweep on 2 ranks
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -4
( timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -s --timeout 800 ) > gpurun_out/n2_rank_parity.log 2>&1
grep -E "RANK_PARITY|passed|failed" gpurun_out/n2_rank_parity.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/n2_bench_cfg4.json 2> gpurun_out/n2_bench_cfg4.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --workload cfg5 --inference --batch 4096 > gpurun_out/n2_bench_cfg5_inference.json 2> gpurun_out/n2_bench_cfg5_inference.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/n2_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("n2_bench_")[1], j["n_gpus"], round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), j["loss_final"], j["clocks"]["sm_mhz"], {k:v["ms"] for k,v in j["kernel_breakdown_ms"].items() if isinstance(v,dict)}, "step", round(r["step_hbm_frac"],3))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
