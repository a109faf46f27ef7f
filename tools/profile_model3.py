#!/usr/bin/env python
"""torch.profiler view of one model-#3 training step (where the time outside libdadmm goes)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
import gnn_dlasso_models_progressive as M, gnn_dlasso_utils
w = dict(P=5, n=500, m=100, K=15, B=1024, graph_prob=0.5)
dev = torch.device("cuda:0")
args, A, label, graphs, _ = bench.make_problem(w, w["B"])
args.GHyp_hidden = 100
label = label.to(dev)
b = torch.stack([A[0, p].to(dev) @ label for p in range(w["P"])], dim=1).contiguous()
torch.manual_seed(0)
model = M.DLASSO_GNNHyp3_Progressive(A, args).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
def step():
    Y, hyp = model(b, graphs, training_iterations=15)
    lm, lf = gnn_dlasso_utils.compute_loss(Y, label, check_finite=False)
    opt.zero_grad(); lf.backward(); torch.nn.utils.clip_grad_norm_(model.parameters(), 100.0); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
ka = prof.key_averages()
cuda_total = sum(getattr(e, "self_device_time_total", getattr(e, "self_cuda_time_total", 0)) for e in ka) / 1e3
print(f"sum of GPU kernel time: {cuda_total:.1f} ms; ops recorded: {sum(e.count for e in ka)}")
print(ka.table(sort_by="self_cuda_time_total", row_limit=18, max_name_column_width=60))
print(ka.table(sort_by="self_cpu_time_total", row_limit=12, max_name_column_width=60))
# kernels only, grouped by name (the op rows above double-count: an op's row includes its kernels)
from collections import defaultdict
from torch.autograd import DeviceType
acc = defaultdict(lambda: [0.0, 0])
for e in prof.events():
    if e.device_type == DeviceType.CUDA:
        dur = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0) or (e.time_range.end - e.time_range.start)
        acc[e.name][0] += dur / 1e3
        acc[e.name][1] += 1
tot = sum(v[0] for v in acc.values())
print(f"GPU kernels only: {tot:.2f} ms in {sum(v[1] for v in acc.values())} launches")
for name, (ms, cnt) in sorted(acc.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{ms:8.3f} ms {cnt:5d} x {ms / cnt * 1e3:7.1f} us  {name[:150]}")
