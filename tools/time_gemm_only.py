#!/usr/bin/env python
"""Kernel-only time of the contraction kinds (profiler events), cfg4 shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hyperparameter-gnn_unfolded-d-admm-main_b200"))
import torch
from dadmm_b200 import functional as DF, _lib
algo = sys.argv[1] if len(sys.argv) > 1 else "f16"
acc = len(sys.argv) > 2 and sys.argv[2] == "acc"
B, P, n = 4096, 50, 1024
W = torch.randn(P, n, n, device="cuda"); x = torch.randn(B, P, n, device="cuda"); out = torch.empty_like(x)
for _ in range(3): DF.contract(W, x, out=out, algo=algo, accumulate=acc)
torch.cuda.synchronize(); _lib.profile_enable(True)
for _ in range(10): DF.contract(W, x, out=out, algo=algo, accumulate=acc)
torch.cuda.synchronize(); pr = _lib.profile_read(); _lib.profile_enable(False)
print(algo, 'accumulate' if acc else 'plain', {k: round(v[0] / max(v[1], 1), 3) for k, v in pr.items() if v[1]})
