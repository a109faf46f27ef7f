#!/usr/bin/env python
"""K-step parity at the bench workload's own P, n, m, K (reduced batch): every contraction route against the library's
fp64 instantiation on the same inputs -- Y[k] rel-L2 and loss_final.  The gate of tests/test_gpu_parity.py
(err <= max(1e-5, 2 x the exact-FMA fp32 path's own distance to fp64)) at full problem size.

    python tools/parity_cfg4.py [--workload cfg4] [--batch 256]
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
ap = argparse.ArgumentParser(); ap.add_argument("--workload", default="cfg4"); ap.add_argument("--batch", type=int, default=256)
o = ap.parse_args()
import unfolded_DLASSO, gnn_dlasso_utils
w = dict(bench.WORKLOADS[o.workload]); B = o.batch; w["B"] = B
dev = torch.device("cuda:0")
args, A, label, graphs, param = bench.make_problem(w, B)
b = torch.stack([A[0, p] @ label for p in range(w["P"])], dim=1).contiguous()

def run(dtype, algo, two_stage):
    # the reference's fp64 protocol (oracle/ref_harness.py): default dtype = run dtype, noise drawn in fp32 and cast
    prev, old = torch.get_default_dtype(), torch.randn
    def randn32(*a, **k):
        k.pop("dtype", None)
        return old(*a, **k, dtype=torch.float32).to(dtype)
    torch.set_default_dtype(dtype)
    torch.randn = randn32
    try:
        m = unfolded_DLASSO.DLASSO_unfolded(A.to(dev, dtype), args).to(dev)
        m.contract_algo, m.two_stage = algo, two_stage
        with torch.no_grad():
            m.seq_hyp.param.copy_(param.to(dtype))
        torch.manual_seed(7)
        with torch.no_grad():
            Y, _ = m(b.to(dev, dtype), graphs)
            lm, lf = gnn_dlasso_utils.compute_loss(Y, label.to(dev, dtype), check_finite=False)
    finally:
        torch.randn = old
        torch.set_default_dtype(prev)
    return Y.double().cpu(), float(lf)

Y64, l64 = run(torch.float64, "simt", False)
rel = lambda a, r: float((a - r).norm() / r.norm())
K = w["K"]
res = {}
for tag, algo, two in (("fp32 exact-FMA", "simt", False), ("f16x3 AtA", "f16", False), ("f16x3 two-stage", "f16", True), ("3xTF32 AtA", "tc", False)):
    Y, lf = run(torch.float32, algo, two)
    res[tag] = [rel(Y[k], Y64[k]) for k in (0, K // 2, K - 1)]
    print(f"{tag:18s} Y err k=0/{K//2}/{K-1}: {res[tag][0]:.2e} {res[tag][1]:.2e} {res[tag][2]:.2e}   loss_final {lf:.7f} (fp64 {l64:.7f}, rel {abs(lf-l64)/l64:.1e})")
ref = res["fp32 exact-FMA"]
for tag, e in res.items():
    ok = all(x <= max(1e-5, 2 * r) for x, r in zip(e, ref))
    print(f"gate {tag:18s}: {'PASS' if ok else 'FAIL'}")
