#!/usr/bin/env python
"""Per-kind kernel times of one cfg4 step: python tools/time_levels.py [--K 6]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
ap = argparse.ArgumentParser(); ap.add_argument("--K", type=int, default=6); ap.add_argument("--workload", default="cfg4")
o = ap.parse_args()
import unfolded_DLASSO, gnn_dlasso_utils
from dadmm_b200 import _lib
w = dict(bench.WORKLOADS[o.workload]); w["K"] = o.K
dev = torch.device("cuda:0")
args, A, label, graphs, param = bench.make_problem(w, w["B"])
A, label = A.to(dev), label.to(dev)
b = torch.stack([A[0, p] @ label for p in range(w["P"])], dim=1).contiguous()
model = unfolded_DLASSO.DLASSO_unfolded(A, args).to(dev)
with torch.no_grad(): model.seq_hyp.param.copy_(param)
def step():
    Y, _ = model(b, graphs); lm, lf = gnn_dlasso_utils.compute_loss(Y, label, check_finite=False); model.zero_grad(); lf.backward()
step(); step(); torch.cuda.synchronize()
_lib.profile_enable(True); step(); torch.cuda.synchronize(); pr = _lib.profile_read(); _lib.profile_enable(False)
print({k: (round(v[0] / max(v[1], 1), 3), v[1]) for k, v in pr.items() if v[1]})
