#!/usr/bin/env python
"""Model #3 (GNN hypernetwork between iterations), BASELINE configs[1] shapes: P=5, n=500, m=100, batch 1024, K iterations.
    python tools/bench_model3.py [--K 15] [--batch 1024] [--hidden 100]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
ap = argparse.ArgumentParser()
ap.add_argument("--K", type=int, default=15); ap.add_argument("--batch", type=int, default=1024); ap.add_argument("--hidden", type=int, default=100)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--cuda-graph", action="store_true", help="capture the whole training step into one CUDA graph (dadmm_b200/graphs.py)")
o = ap.parse_args()
import gnn_dlasso_models_progressive as M, gnn_dlasso_utils
from dadmm_b200 import _lib
w = dict(P=5, n=500, m=100, K=o.K, B=o.batch, graph_prob=0.5)
dev = torch.device("cuda:0")
args, A, label, graphs, _ = bench.make_problem(w, o.batch)
args.GHyp_hidden = o.hidden
label = label.to(dev)
b = torch.stack([A[0, p].to(dev) @ label for p in range(w["P"])], dim=1).contiguous()
torch.manual_seed(0)
model = M.DLASSO_GNNHyp3_Progressive(A, args).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, capturable=o.cuda_graph)
def eager_step(bb=None, ll=None):
    Y, hyp = model(b if bb is None else bb, graphs, training_iterations=o.K)
    lm, lf = gnn_dlasso_utils.compute_loss(Y, label if ll is None else ll, check_finite=False)
    opt.zero_grad(); lf.backward(); torch.nn.utils.clip_grad_norm_(model.parameters(), 100.0); opt.step()
    return lf.detach()
step = eager_step
if o.cuda_graph:
    from dadmm_b200.graphs import GraphedStep
    model.check_finite = "deferred"
    graphed = GraphedStep(eager_step, [b, label], warmup=3)
    step = lambda: graphed(b, label)
for _ in range(2): step()
torch.cuda.synchronize(); n0 = _lib.launch_count(); t0 = time.perf_counter()
for _ in range(o.steps): lf = step()
torch.cuda.synchronize(); t = (time.perf_counter() - t0) / o.steps
print(f"model3 P=5 n=500 B={o.batch} K={o.K} hidden={o.hidden}: {1e3*t:.1f} ms/step, {o.K*o.batch/t:.0f} iter*problems/s, "
      f"loss_final={float(lf.detach()):.5f}, dadmm kernels/step={(_lib.launch_count()-n0)//o.steps}" + (" [one CUDA graph per step]" if o.cuda_graph else ""))
ms_step, loss_val, per_step = 1e3 * t, float(lf.detach()), (_lib.launch_count() - n0) // o.steps
_lib.profile_enable(True); torch.cuda.synchronize(); t0 = time.perf_counter(); eager_step(); torch.cuda.synchronize(); t1 = time.perf_counter() - t0
pr = _lib.profile_read(); _lib.profile_enable(False)
lib_ms = sum(v[0] for v in pr.values())
print(f"profiled step {1e3*t1:.1f} ms; libdadmm kernels {lib_ms:.1f} ms:", {k: (round(v[0], 2), v[1]) for k, v in pr.items() if v[1]})

# one JSON line in bench.py's format (side measurement: BASELINE configs[1] is not the metric's config)
print(json.dumps({"metric": "unfolded D-ADMM iterations*problems/sec (fwd+bwd)", "value": o.K * o.batch / (ms_step / 1e3), "unit": "iter*problems/s",
                  "n_gpus": 1, "steps": o.steps, "warmup": 2, "ms_per_step": ms_step, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
                  "config": {"workload": "BASELINE configs[1]: model #3 (GNN hypernetwork between iterations), P=5, n=500, m=100, "
                                         f"batch {o.batch}, K={o.K}, hidden {o.hidden}, fresh ER p=0.5 graph per problem",
                             "step": "forward K iterations + compute_loss + backward + clip_grad_norm_ + AdamW",
                             "launch": "whole step replayed as one CUDA graph" if o.cuda_graph else "eager"},
                  "loss_final": loss_val, "libdadmm_ms_in_profiled_step": round(lib_ms, 2),
                  "libdadmm_launches_per_step": per_step if not o.cuda_graph else "captured"}))
