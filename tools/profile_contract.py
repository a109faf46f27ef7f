#!/usr/bin/env python
"""Contraction only, config-4 shapes: python tools/profile_contract.py [--B 4096] [--P 50] [--n 1024] [--reps 3] [--algo tc]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hyperparameter-gnn_unfolded-d-admm-main_b200"))
import torch
from dadmm_b200 import functional as DF

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=4096); ap.add_argument("--P", type=int, default=50)
ap.add_argument("--n", type=int, default=1024); ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--algo", default="tc")
o = ap.parse_args()
W = torch.randn(o.P, o.n, o.n, device="cuda"); x = torch.randn(o.B, o.P, o.n, device="cuda")
out = torch.empty_like(x)
DF.contract(W, x, out=out, algo=o.algo)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(o.reps):
    DF.contract(W, x, out=out, algo=o.algo)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / o.reps
print(f"algo={o.algo} variant={os.environ.get('DADMM_TC_VARIANT','2')} B={o.B} P={o.P} n={o.n}: {ms:.3f} ms/launch, {2*o.P*o.n*o.n*o.B/ms/1e9:.1f} TFLOP/s (fp32-equivalent)")
