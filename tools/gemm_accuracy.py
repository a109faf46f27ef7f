#!/usr/bin/env python
"""Accuracy (rel-L2 vs fp64) and kernel time of the tensor-core contraction kinds at cfg4-like shapes.
    DADMM_F16_KBC=1|2|4 python tools/gemm_accuracy.py      (k-blocks of 64 per tensor-core partial sum)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hyperparameter-gnn_unfolded-d-admm-main_b200"))
import torch
from dadmm_b200 import functional as DF, _lib
torch.manual_seed(0)
B, P = 1024, 4
for n_out, n_in in ((1024, 1024), (256, 1024), (1024, 256), (2048, 2048)):
    W = torch.randn(P, n_out, n_in, device="cuda") / n_in ** 0.5
    x = torch.randn(B, P, n_in, device="cuda") * (torch.rand(B, P, n_in, device="cuda") < 0.5)
    ref = torch.einsum("pik,bpk->bpi", W.double(), x.double())
    res = {}
    for algo in ("simt", "tc", "f16"):
        o = DF.contract(W, x, algo=algo)
        res[algo] = float((o.double() - ref).norm() / ref.norm())
    print(f"kbc={os.environ.get('DADMM_F16_KBC', 'default')} n_out={n_out} n_in={n_in}: " + "  ".join(f"{k}={v:.2e}" for k, v in res.items()))
