"""CPU probe: does a two-stage contraction A^T (A y) -- half the flops of AtA y at m = n/4 -- keep the K-step
parity gate?  Variants are run through the oracle recurrence and compared with the reference's fp64 run.

TEST INFRASTRUCTURE (lives under tests/ because it executes the oracle); not collected by pytest, run by hand:
    python tests/probes/two_stage_probe.py"""
import sys, os
TESTS = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(TESTS))
sys.path.insert(0, TESTS)
import torch
from helpers import Golden, rel_l2
from oracle import dadmm_oracle as O

def split16(x):
    """scaled fp16 hi/lo split of a fp32 tensor (per-tensor power-of-two scale), returned as fp64 hi+lo value"""
    amax = x.abs().max().item()
    import math
    e = 13 - math.floor(math.log2(amax)) if amax > 0 else 0
    s = 2.0 ** e
    xs = x.double() * s
    hi = xs.float().half()
    lo = (xs - hi.double()).float().half()
    return hi.double() / s, lo.double() / s

def mm3(Whi, Wlo, xhi, xlo, eq):
    return torch.einsum(eq, Whi, xhi) + torch.einsum(eq, Whi, xlo) + torch.einsum(eq, Wlo, xhi)

def run(name, g, contract_fn, grad=True):
    A = g.t("A"); b = g.t("b"); label = g.t("label")
    param = g.t("param").clone().requires_grad_(True)
    table = O.hyp_table(param, g.t("max_param").float(), training=bool(g.z["training"]))
    y0, U0, d0 = g.t("y0"), g.t("U0"), g.t("d0")
    Atb = O.atx(A, b)
    deg = O.degrees(g.graphs, g.P)
    lap2 = O.laplacian2(g.graphs, g.P)
    y, U, d = y0, U0, d0
    Y = []
    for k in range(g.K):
        h = table[k]
        al, ta, rh, et = (h[:, i].reshape(1, -1, 1, 1).expand(1, g.P, 1, 1) for i in range(4))   # 'same' tables have one row
        a = contract_fn(y)
        y, U, d, _ = O.step(a, Atb, deg, y, U, d, al, ta, rh, et, O.clamps_model1(k), lambda v: O.delta_dense(lap2, v))
        Y.append(y)
    Y = torch.stack(Y)
    _, lf = O.loss(Y, label)
    lf.backward()
    Y64 = g.t("Y64")
    errs = [rel_l2(Y[k], Y64[k]) for k in (0, g.K // 2, g.K - 1)]
    print(f"{name:28s} Y err k=0/mid/last: {errs[0]:.2e} {errs[1]:.2e} {errs[2]:.2e}   dparam err {rel_l2(param.grad, g.t('dparam64')):.2e}")

for case in ("m1_trained_P5_n500", "m1_trained15_P5_n51", "m1_zero_P5_n64", "m1_same_pergraph_P8_n48"):
    g = Golden(case)
    print("==", case, "K", g.K)
    A = g.t("A")[0]                  # [P,m,n]
    AtA = torch.einsum("pmi,pmj->pij", A, A)           # as the reference builds it (fp32)
    AtA_ref = O.atx(g.t("A"), g.t("A"))[0] if False else AtA
    A64 = A.double()
    Ahi, Alo = split16(A)
    Whi, Wlo = split16(AtA)
    ref32 = rel_l2(g.t("Y")[-1], g.t("Y64")[-1])
    print(f"   reference fp32 vs fp64: Y last {ref32:.2e}  dparam {rel_l2(g.t('dparam'), g.t('dparam64')):.2e}")
    run("ata fp32", g, lambda y: torch.einsum("pij,bpjo->bpio", AtA, y))
    run("two-stage fp32", g, lambda y: torch.einsum("pmi,bpmo->bpio", A, torch.einsum("pmj,bpjo->bpmo", A, y)))
    run("two-stage exact", g, lambda y: torch.einsum("pmi,bpmo->bpio", A64, torch.einsum("pmj,bpjo->bpmo", A64, y.double())).float())
    def ata_emul(y):
        yh, yl = split16(y.detach())
        # differentiable surrogate: value from the emulation, gradient from the fp32 op
        v = mm3(Whi, Wlo, yh, yl, "pij,bpjo->bpio").float()
        s = torch.einsum("pij,bpjo->bpio", AtA, y)
        return s + (v - s).detach()
    run("ata 3xfp16 (shipping)", g, ata_emul)
    def ts_emul(y):
        yh, yl = split16(y.detach())
        t = mm3(Ahi, Alo, yh, yl, "pmj,bpjo->bpmo").float()
        th, tl = split16(t)
        v = mm3(Ahi, Alo, th, tl, "pmi,bpmo->bpio").float()
        s = torch.einsum("pij,bpjo->bpio", AtA, y)
        return s + (v - s).detach()
    run("two-stage 3xfp16", g, ts_emul)
