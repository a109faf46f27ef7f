"""Parity at the BASELINE shapes against the CPU oracle (not against the library's own fp64 build).

The oracle (``oracle/dadmm_oracle.py``) is pinned bit for bit to the unmodified reference on the golden fixtures; those
fixtures are small (P <= 12, n <= 500) and none of them reaches the kernels the bench workload runs -- the two-stage
tcgen05 contraction ``contract_f16_kernel`` (needs B >= 128, n_out > 128), the LEAN level kernels (n % 128 == 0), staged
neighbour lists of 50-node graphs.  These tests run the default route of the drop-in module at the dimensions of
BASELINE.json configs[3] (P=50, n=1024, m=256) and configs[2] (P=20, n=256, m=64) on a reduced batch and compare
trajectory, loss and d loss / d seq_hyp.param with the oracle's fp64 run on the same A, b, graphs, noise and table
(reference lines: unfolded_DLASSO.py:53-107, gnn_dlasso_utils.py:27-88).

Gate: ``err <= max(1e-5, 2 * err_ref32)`` where ``err_ref32`` is the distance of the oracle's own fp32 run (the
reference's arithmetic, bit-faithful accumulation order) from its fp64 run -- SURVEY 8c: with ``set_A``'s conditioning
and an untrained table the map is expanding, so the reference's fp32 run itself leaves 1e-5 after a few iterations.  The
contracting-regime test (A scaled to singular values <= 1, all K = 25 iterations) is where the plain 1e-5 gate is
meaningful, and there it is applied untrimmed.  Every test prints the plain rel-L2 next to the trimmed one.
"""
import argparse
import math
import os
import sys

import pytest
import torch

from helpers import ROOT, rel_l2
from oracle import dadmm_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MAXP = torch.tensor([0.1, 0.99, 0.99, 0.99])


def trimmed_rel_l2(a, r, frac=1e-4):
    """rel-L2 over all but the ``frac`` worst elements (sign / clamp switching points; printed beside the plain norm)."""
    d = (a.double() - r.double()).abs().flatten()
    k = max(1, int(d.numel() * (1.0 - frac)))
    thr = d.kthvalue(k).values
    return float((d[d <= thr] ** 2).sum().sqrt() / r.double().norm())


def _problem(workload, B, K, a_scale=1.0, param_scale=1.0):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import bench
    w = dict(bench.WORKLOADS[workload])
    w["K"], w["B"] = K, B
    args, A, label, graphs, param = bench.make_problem(w, B, set_A=O.set_A)
    A = A * a_scale
    b = torch.stack([A[0, p] @ label for p in range(w["P"])], dim=1).contiguous()
    gen = torch.Generator().manual_seed(11)
    noise = [torch.randn((B, w["P"], w["n"], 1), generator=gen) for _ in range(3)]       # the three N(0,1) draws (:49-51)
    return w, args, A, b, label, graphs, param[:K] * param_scale, noise


def _oracle(A, b, label, graphs, param, noise, dtype, exact):
    """Oracle run in ``dtype``: (Y [K,B,P,n,1], loss_final, d loss_final / d param) -- gradient only for the dense-2L form."""
    A, b, label = A.to(dtype), b.to(dtype), label.to(dtype)
    y0, U0, d0 = (t.to(dtype) * 1e-2 for t in noise)
    prm = param.to(dtype).clone().requires_grad_(not exact)
    table = O.hyp_table(prm, MAXP.to(dtype), training=True)
    P = A.shape[1]
    hyp = table.expand(table.shape[0], P, 4)
    with torch.set_grad_enabled(not exact):
        # exact: the reference's arithmetic bit for bit (per-agent mat-vec loops, accumulation order of compute_delta);
        # otherwise the differentiable form with batched GEMMs (same values up to summation order)
        Y = O.unfolded_forward(O.atx(A, A), O.atx(A, b), graphs, y0, U0, d0, hyp, exact_delta=exact, gemm_contract=not exact)
        _, lf = O.loss(Y, label, vectorised=not exact)
    grad = None
    if not exact:
        lf.backward()
        grad = prm.grad.detach()
    return Y.detach(), float(lf.detach()), grad


def _oracle_forward(A, b, label, graphs, param, noise, dtype, exact):
    """Forward-only oracle run (no autograd graph: the configs[4] dimensions would not fit one)."""
    A, b = A.to(dtype), b.to(dtype)
    y0, U0, d0 = (t.to(dtype) * 1e-2 for t in noise)
    table = O.hyp_table(param.to(dtype), MAXP.to(dtype), training=True)
    hyp = table.expand(table.shape[0], A.shape[1], 4)
    return O.unfolded_forward(O.atx(A, A), O.atx(A, b), graphs, y0, U0, d0, hyp, exact_delta=exact, gemm_contract=not exact)


def _ours(args, A, b, label, graphs, param, noise, algo="auto", two_stage=True):
    import unfolded_DLASSO
    import gnn_dlasso_utils
    model = unfolded_DLASSO.DLASSO_unfolded(A.to(DEV), args).to(DEV)
    model.contract_algo, model.two_stage = algo, two_stage
    with torch.no_grad():
        model.seq_hyp.param.copy_(param)
    Y, _ = model(b.to(DEV), graphs, noise=[t.to(DEV) * 1e-2 for t in noise])
    _, lf = gnn_dlasso_utils.compute_loss(Y, label.to(DEV))
    lf.backward()
    return Y.detach().cpu(), float(lf.detach()), model.seq_hyp.param.grad.detach().cpu()


def _report(tag, Y, Y32, Y64, g, g32, g64, lf, l32, l64):
    K = Y.shape[0]
    plain = [rel_l2(Y[k], Y64[k]) for k in range(K)]
    trim = [trimmed_rel_l2(Y[k], Y64[k]) for k in range(K)]
    ref = [rel_l2(Y32[k], Y64[k]) for k in range(K)]
    show = lambda v: " ".join(f"{x:.1e}" for x in (v if len(v) <= 8 else v[:3] + v[-3:]))
    print(f"\n{tag}: Y[k] vs oracle fp64 rel-L2 plain [{show(plain)}] trimmed(1e-4) [{show(trim)}]; oracle fp32 vs fp64 [{show(ref)}]; "
          f"dparam {rel_l2(g, g64):.2e} (oracle fp32 {rel_l2(g32, g64):.2e}); loss {lf:.7f} / fp32 {l32:.7f} / fp64 {l64:.7f}")
    return plain, trim, ref


def test_configs3_dimensions_default_route_vs_oracle():
    """BASELINE configs[3] dimensions (P=50, n=1024, m=256, ER p=0.12 bridged graphs; 128 problems, K=5, the bench's own
    operator, table and graphs): two-stage ``contract_f16_kernel`` + LEAN level kernels vs the oracle."""
    from dadmm_b200 import _lib
    w, args, A, b, label, graphs, param, noise = _problem("cfg4", B=128, K=5)
    assert _lib.lib.dadmm_unfolded_uses_factor(_lib.F32, _lib.ALGO_AUTO, 128, w["P"], w["n"], w["m"]) == 1
    Y64, l64, g64 = _oracle(A, b, label, graphs, param, noise, torch.float64, exact=False)
    Y32, l32, _ = _oracle(A, b, label, graphs, param, noise, torch.float32, exact=True)      # the reference's fp32 arithmetic
    _, _, g32 = _oracle(A, b, label, graphs, param, noise, torch.float32, exact=False)
    Y, lf, g = _ours(args, A, b, label, graphs, param, noise)
    plain, trim, ref = _report("configs[3] dims, default route", Y, Y32, Y64, g, g32, g64, lf, l32, l64)
    for k in range(w["K"]):
        assert plain[k] <= max(1e-5, 2 * ref[k]), (k, plain[k], ref[k])
    assert abs(lf - l64) <= max(1e-5, 2 * abs(l32 - l64) / abs(l64)) * abs(l64)
    assert rel_l2(g, g64) <= max(1e-5, 2 * rel_l2(g32, g64)), (rel_l2(g, g64), rel_l2(g32, g64))


def test_configs2_dimensions_default_route_vs_oracle():
    """BASELINE configs[2] dimensions (P=20, n=256, m=64, ER p=0.5; 256 problems, K=8): single-stage ``contract_f16_kernel``
    (m = 64 is below the two-stage tile) + LEAN level kernels with the label-free loss sums, vs the oracle."""
    from dadmm_b200 import _lib
    w, args, A, b, label, graphs, param, noise = _problem("cfg3", B=256, K=8)
    assert _lib.lib.dadmm_contract_uses_tensor_cores(_lib.F32, _lib.ALGO_AUTO, 256, w["P"], w["n"], w["n"]) > 0
    Y64, l64, g64 = _oracle(A, b, label, graphs, param, noise, torch.float64, exact=False)
    Y32, l32, _ = _oracle(A, b, label, graphs, param, noise, torch.float32, exact=True)
    _, _, g32 = _oracle(A, b, label, graphs, param, noise, torch.float32, exact=False)
    Y, lf, g = _ours(args, A, b, label, graphs, param, noise)
    plain, trim, ref = _report("configs[2] dims, default route", Y, Y32, Y64, g, g32, g64, lf, l32, l64)
    for k in range(w["K"]):
        assert plain[k] <= max(1e-5, 2 * ref[k]), (k, plain[k], ref[k])
    assert abs(lf - l64) <= max(1e-5, 2 * abs(l32 - l64) / abs(l64)) * abs(l64)
    assert rel_l2(g, g64) <= max(1e-5, 2 * rel_l2(g32, g64)), (rel_l2(g, g64), rel_l2(g32, g64))


def test_configs3_dimensions_full_K_in_a_contracting_regime():
    """All K = 25 iterations at configs[3] dimensions in a regime where the recurrence contracts: A scaled to singular values
    <= 1 and a trained-style table (alpha ~ 0.03, tau ~ 0.12, rho, eta ~ 0.02: with 50 agents the consensus terms -- not
    AtA -- are what makes an untrained table expanding, alpha * rho * lambda_max(2L) ~ 4).  There fp32 and fp64 stay together
    and the north star's 1e-5 is checkable as it stands, untrimmed, on the loss (final and mean over all K iterations) and
    on d loss / d param.  The iterates carry one more effect the reference's own fp32 run shows identically (measured with
    the oracle: 1 element of 6.5 M off by > 1e-4 at k = 3, ~900 at k = 24, plain rel-L2 1e-4 while the trimmed norm stays
    below 1e-5): ``sign(y) * tau`` is discontinuous, and an element within rounding of zero flips it -- so the iterates are
    gated untrimmed against twice the reference's own fp32 distance and trimmed (1e-4 of the elements) against 1e-5."""
    import gnn_dlasso_utils
    w, args, A, b, label, graphs, _, noise = _problem("cfg4", B=128, K=25, a_scale=0.1)
    gen = torch.Generator().manual_seed(3)
    param = torch.randn((25, w["P"], 4), generator=gen) * 0.05
    param[0] += torch.tensor([-1.0, -2.0, -4.0, -4.0])
    Y64, l64, g64 = _oracle(A, b, label, graphs, param, noise, torch.float64, exact=False)
    Y32, l32, g32 = _oracle(A, b, label, graphs, param, noise, torch.float32, exact=False)
    Y, lf, g = _ours(args, A, b, label, graphs, param, noise)
    plain, trim, ref = _report("configs[3] dims, K=25, contracting", Y, Y32, Y64, g, g32, g64, lf, l32, l64)
    ref_trim = [trimmed_rel_l2(Y32[k], Y64[k]) for k in range(25)]
    off = lambda a: int(((a[-1].double() - Y64[-1]).abs() > 1e-4).sum())
    print(f"elements of Y[24] off by more than 1e-4: ours {off(Y)}, oracle fp32 {off(Y32)} of {Y64[-1].numel()}")
    assert abs(lf - l64) <= 1e-5 * abs(l64), (lf, l64)
    lm64 = float(O.loss(Y64, label.double(), vectorised=True)[0])
    lm = float(gnn_dlasso_utils.compute_loss(Y.to(DEV), label.to(DEV))[0])
    assert abs(lm - lm64) <= 1e-5 * abs(lm64), (lm, lm64)
    assert rel_l2(g, g64) <= 1e-5, rel_l2(g, g64)
    for k in range(25):
        assert plain[k] <= max(1e-5, 2 * ref[k]), (k, plain[k], ref[k])
        assert trim[k] <= max(1e-5, 2 * ref_trim[k]), (k, trim[k], ref_trim[k])


def test_configs4_dimensions_inference_vs_oracle():
    """BASELINE configs[4] dimensions (P=100, n=2048, m=512, ER p=0.1 bridged; 128 problems, K=3) on the inference path
    (``no_grad``: no saved streams, U ping-pong, the forward level's one-tile geometry at P=100, two-stage contraction with
    m=512) against the oracle's fp64 run, with the oracle's bit-faithful fp32 run as the yardstick."""
    import unfolded_DLASSO
    import gnn_dlasso_utils
    w, args, A, b, label, graphs, param, noise = _problem("cfg5", B=128, K=3)
    with torch.no_grad():
        Y64 = _oracle_forward(A, b, label, graphs, param, noise, torch.float64, exact=False)
        Y32 = _oracle_forward(A, b, label, graphs, param, noise, torch.float32, exact=True)
        model = unfolded_DLASSO.DLASSO_unfolded(A.to(DEV), args).to(DEV)
        model.seq_hyp.param.copy_(param)
        Y, _ = model(b.to(DEV), graphs, noise=[t.to(DEV) * 1e-2 for t in noise])
        lm, lf = gnn_dlasso_utils.compute_loss(Y, label.to(DEV))
    Y = Y.cpu()
    l64 = float(O.loss(Y64, label.double(), vectorised=True)[1])
    plain = [rel_l2(Y[k], Y64[k]) for k in range(w["K"])]
    ref = [rel_l2(Y32[k], Y64[k]) for k in range(w["K"])]
    print(f"\nconfigs[4] dims, inference: Y[k] vs oracle fp64 rel-L2 {plain}; oracle fp32 vs fp64 {ref}; loss {float(lf):.7f} vs {l64:.7f}")
    for k in range(w["K"]):
        assert plain[k] <= max(1e-5, 2 * ref[k]), (k, plain[k], ref[k])
    assert abs(float(lf) - l64) <= 1e-5 * abs(l64)
