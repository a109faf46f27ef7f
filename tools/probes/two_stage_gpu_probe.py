"""Per-iteration error of every contraction route vs the fp64 instantiation at cfg4 dimensions (small batch)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, bench
from helpers import rel_l2
import unfolded_DLASSO, gnn_dlasso_utils
DEV = "cuda:0"
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
w = dict(bench.WORKLOADS["cfg4"]); w["K"], w["B"] = 6, 128
args, A, label, graphs, param = bench.make_problem(w, w["B"])
A = A * scale
b = torch.stack([A[0, p] @ label for p in range(w["P"])], dim=1).contiguous()
args.GHN_iter_num = w["K"]
def run(dtype, algo, two_stage, rhs=True):
    prev, old = torch.get_default_dtype(), torch.randn
    def randn32(*a, **k):
        k.pop("dtype", None); return old(*a, **k, dtype=torch.float32).to(dtype)
    torch.set_default_dtype(dtype); torch.randn = randn32
    try:
        m = unfolded_DLASSO.DLASSO_unfolded(A.to(DEV, dtype), args).to(DEV)
        m.contract_algo, m.two_stage, m.two_stage_rhs = algo, two_stage, rhs
        with torch.no_grad(): m.seq_hyp.param.copy_(param[: w["K"]].to(dtype))
        torch.manual_seed(7)
        with torch.no_grad(): Y, _ = m(b.to(DEV, dtype), graphs)
    finally:
        torch.randn = old; torch.set_default_dtype(prev)
    return Y.double().cpu()
Y64 = run(torch.float64, "simt", False)
for tag, a in (("exact-FMA", ("simt", False)), ("f16 single", ("f16", False)), ("f16 two-stage+rhs", ("f16", True, True)),
               ("f16 two-stage, Atb in stage 2", ("f16", True, False)), ("3xTF32", ("tc", False))):
    Y = run(torch.float32, *a)
    print(f"{tag:32s}", " ".join(f"{rel_l2(Y[k], Y64[k]):.2e}" for k in range(w["K"])))
