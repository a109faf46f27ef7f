"""``dadmm_gcn_epilogue_fwd/bwd`` (the per-problem part of a graph-convolution layer of the model-#3 hypernetwork, reference
gnn_dlasso_models_progressive.py:37-72) against the PyTorch composition it replaces -- which the CPU suite pins to the
reference's per-sample loop (tests/test_cpu_host.py) -- forward and every gradient, training and eval statistics, with and
without a dropout mask; then the whole encoder, fused against composed."""
import networkx as nx
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _adj(B, P, seed):
    import gnn_dlasso_models_progressive as M
    graphs = [nx.erdos_renyi_graph(P, 0.5, seed=seed + i) for i in range(B)]
    return M.normalized_adjacency(graphs, P, DEV), graphs


def _composed(H, adj, bias, w, b, rm, rv, training, eps, mask):
    z = torch.baddbmm(bias, adj, H)
    a = F.leaky_relu(z, 0.01)
    if training:
        var, mean = torch.var_mean(a, dim=1, unbiased=False, keepdim=True)
        out = (a - mean) * torch.rsqrt(var + eps) * w + b
    else:
        mean = var = None
        out = (a - rm) * torch.rsqrt(rv + eps) * w + b
    if mask is not None:
        out = out * mask
    return out, mean, var


@pytest.mark.parametrize("B,P,C", [(64, 5, 100), (33, 20, 400), (7, 50, 130)])
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("with_mask", [False, True])
def test_epilogue_matches_the_composed_ops(B, P, C, training, with_mask):
    from dadmm_b200 import functional as DF
    gen = torch.Generator(device=DEV).manual_seed(B * 131 + P)
    adj, _ = _adj(B, P, seed=B)
    rnd = lambda *s: torch.randn(*s, device=DEV, generator=gen)
    H = rnd(B, P, C)
    bias, w, b = rnd(C) * 0.3, 1 + 0.2 * rnd(C), 0.1 * rnd(C)
    rm, rv = 0.2 * rnd(C), 0.5 + torch.rand(C, device=DEV, generator=gen)
    mask = (torch.rand(B, P, C, device=DEV, generator=gen) < 0.9).float() / 0.9 if with_mask else None
    gout = rnd(B, P, C)
    outs = []
    for fused in (True, False):
        leaves = [t.clone().requires_grad_(True) for t in (H, bias, w, b)]
        if fused:
            out, mean, var = DF.GCNEpilogue.apply(leaves[0], adj, leaves[1], leaves[2], leaves[3], rm, rv, training, 1e-5, 0.01, mask)
        else:
            out, mean, var = _composed(leaves[0], adj, leaves[1], leaves[2], leaves[3], rm, rv, training, 1e-5, mask)
        (out * gout).sum().backward()
        outs.append((out.detach(), mean, var, [t.grad for t in leaves]))
    (o1, m1, v1, g1), (o2, m2, v2, g2) = outs
    assert rel_l2(o1, o2) < 2e-6
    if training:
        assert rel_l2(m1, m2.squeeze(1)) < 2e-6 and rel_l2(v1, v2.squeeze(1)) < 1e-5
    for name, a, r in zip(("H", "bias", "bn_w", "bn_b"), g1, g2):
        assert rel_l2(a, r) < 2e-5, (name, rel_l2(a, r))


def test_encoder_fused_equals_composed_in_training_mode():
    """GNNHypernetwork3 with the kernel path against the composed PyTorch path: same output, same parameter gradients, same
    running statistics after one training-mode pass (dropout switched off so that both see the same activations).  Five
    layers of BatchNorm over P = 5 samples amplify fp32 rounding (1.5e-5 between the two fp32 evaluations), so both are
    measured against the composed path in float64: the kernel path may be no further from it than twice the composed fp32
    path is."""
    import copy
    import gnn_dlasso_models_progressive as M
    B, P, m, hidden = 96, 5, 64, 24
    torch.manual_seed(3)
    enc = M.GNNHypernetwork3(P, m, hidden).to(DEV).train()
    enc.dropout.p = 0.0
    ref = copy.deepcopy(enc)
    ref.fused = False
    ref64 = copy.deepcopy(enc).double()
    ref64.fused = False
    adj, graphs = _adj(B, P, seed=11)
    x = torch.randn(B, P, m, 1, device=DEV)
    gy = torch.randn(B, P * 4 * hidden, device=DEV)
    res = []
    for net, dt in ((enc, torch.float32), (ref, torch.float32), (ref64, torch.float64)):
        xi = x.to(dt).detach().clone().requires_grad_(True)
        y = net(xi, graphs, adj.to(dt))
        (y * gy.to(dt)).sum().backward()
        res.append((y.detach(), xi.grad, {k: p.grad for k, p in net.named_parameters()}, {k: v.clone() for k, v in net.named_buffers()}))
    (y1, gx1, gp1, bf1), (y2, gx2, gp2, bf2), (y3, gx3, gp3, bf3) = res
    assert rel_l2(y1, y3) <= 2 * rel_l2(y2, y3) + 1e-6, (rel_l2(y1, y3), rel_l2(y2, y3))
    assert rel_l2(gx1, gx3) <= 2 * rel_l2(gx2, gx3) + 1e-5, (rel_l2(gx1, gx3), rel_l2(gx2, gx3))
    for k in gp3:
        assert rel_l2(gp1[k], gp3[k]) <= 2 * rel_l2(gp2[k], gp3[k]) + 2e-5, (k, rel_l2(gp1[k], gp3[k]), rel_l2(gp2[k], gp3[k]))
    for k in bf3:
        assert rel_l2(bf1[k].double(), bf3[k].double()) < 1e-5, k
    print(f"encoder output vs float64: kernel path {rel_l2(y1, y3):.2e}, composed fp32 {rel_l2(y2, y3):.2e}")


def test_set_Data_on_the_device_matches_the_reference_loop_and_rng_position():
    """``gnn_data.set_Data`` with ``A`` on the GPU forms all P observation vectors in one contraction launch
    (zero agent stride for the shared label) -- same values as the reference's per-agent matmul loop (gnn_data.py:13-14) to
    fp32 rounding, same random stream position afterwards."""
    import argparse
    import gnn_data
    P, m, n, N = 7, 24, 96, 50
    A = torch.randn(1, P, m, n, device=DEV)
    torch.manual_seed(5)
    b, y = gnn_data.make_problem(A, N)
    after = torch.randn(4, device=DEV)
    torch.manual_seed(5)
    y2 = 2 * torch.randn(N, n, 1, device=DEV) * (torch.rand(N, n, 1, device=DEV) <= 0.25)
    b2 = torch.randn(N, P, m, 1, device=DEV)
    for p in range(P):
        b2[:, p] = torch.matmul(A[0, p], y2)
    assert torch.equal(y, y2) and torch.equal(after, torch.randn(4, device=DEV))
    assert rel_l2(b, b2) < 1e-6
    loader = gnn_data.set_Data(A, N, argparse.Namespace(batch_size=16, snr=4))
    bb, yy = next(iter(loader))
    assert bb.shape == (16, P, m, 1) and yy.shape == (16, n, 1) and bb.is_cuda


@pytest.mark.parametrize("M,c_in,c_out", [(5120, 400, 400), (5120, 100, 200), (1024, 2000, 400), (5120, 1000, 100), (96, 64, 24)])
def test_tensor_core_linear_matches_float64(M, c_in, c_out):
    """``DF.linear`` (hypernetwork products on the fp32-accurate tensor-core contraction where the shape takes it, PyTorch's
    own otherwise): output, input gradient and weight gradient against float64, no further off than twice cuBLAS fp32."""
    from dadmm_b200 import functional as DF
    gen = torch.Generator(device=DEV).manual_seed(M + c_in)
    x = torch.randn(M, c_in, device=DEV, generator=gen)
    W = torch.randn(c_out, c_in, device=DEV, generator=gen) / c_in ** 0.5
    b = torch.randn(c_out, device=DEV, generator=gen)
    g = torch.randn(M, c_out, device=DEV, generator=gen)
    res = []
    for fn, dt in ((DF.linear, torch.float32), (F.linear, torch.float32), (F.linear, torch.float64)):
        xs, Ws, bs = (t.to(dt).clone().requires_grad_(True) for t in (x, W, b))
        y = fn(xs, Ws, bs)
        (y * g.to(dt)).sum().backward()
        res.append((y.detach(), xs.grad, Ws.grad, bs.grad))
    ours, ref32, ref64 = res
    for name, a, r32, r64 in zip(("y", "dx", "dW", "db"), ours, ref32, ref64):
        assert rel_l2(a, r64) <= 2 * rel_l2(r32, r64) + 2e-7, (name, rel_l2(a, r64), rel_l2(r32, r64))
