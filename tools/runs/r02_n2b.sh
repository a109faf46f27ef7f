# round 2, second 2-GPU call (final build): same-noise rank parity and the strong-scaling bench line
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -s --timeout 800 ) > gpurun_out/n2b_rank_parity.log 2>&1
grep -E "RANK_PARITY|passed|failed" gpurun_out/n2b_rank_parity.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/n2b_bench_cfg4.json 2> gpurun_out/n2b_bench_cfg4.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/n2b_bench_cfg4.json").read().strip().splitlines()[-1])
print(j["n_gpus"], round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), j["loss_final"], j["clocks"]["sm_mhz"])
PY
