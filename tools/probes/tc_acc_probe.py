import sys, torch
sys.path.insert(0, "hyperparameter-gnn_unfolded-d-admm-main_b200")
from dadmm_b200 import functional as DF
dev = "cuda:0"
def rel(a, b): return float((a.double()-b.double()).norm()/b.double().norm())
def trunc_tf32(t): return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)
for n in (256, 512, 1024, 2048):
    B, P = 512, 2
    g = torch.Generator().manual_seed(n)
    W = torch.randn((P, n, n), generator=g); x = torch.randn((B, P, n), generator=g)
    import os
    for name, (Wi, xi) in {"full fp32 inputs": (W, x), "tf32-exact inputs": (trunc_tf32(W), trunc_tf32(x)),
                            "positive tf32-exact": (trunc_tf32(W.abs()), trunc_tf32(x.abs()))}.items():
        ref = torch.einsum("pik,bpk->bpi", Wi.double(), xi.double())
        o_tc = DF.contract(Wi.to(dev), xi.to(dev), algo=os.environ.get("PROBE_ALGO","tc")).cpu()
        o_si = DF.contract(Wi.to(dev), xi.to(dev), algo="simt").cpu()
        bias = float(((o_tc.double()-ref)/ref.abs().clamp_min(1e-30)).mean())
        print(f"n={n:5d} {name:22s} tc={rel(o_tc,ref):.2e} simt={rel(o_si,ref):.2e} mean signed rel err tc={bias:+.2e}")
