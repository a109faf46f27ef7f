# round 2, second 4-GPU call (final build): strong-scaling bench line of configs[3]
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/n4b_bench_cfg4.json 2> gpurun_out/n4b_bench_cfg4.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/n4b_bench_cfg4.json").read().strip().splitlines()[-1])
print(j["n_gpus"], round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), j["loss_final"], j["clocks"]["sm_mhz"])
PY
