mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8_v23.json 2> gpurun_out/bench_n8_v23.err
python - <<'PY'
import json
try:
    j=json.loads(open("gpurun_out/bench_n8_v23.json").read().strip().splitlines()[-1])
    print(j["n_gpus"], round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3), j["loss_final"], j["clocks"], j["kernel_breakdown_ms"])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/bench_n8_v23.err").read()[-2500:])
PY
