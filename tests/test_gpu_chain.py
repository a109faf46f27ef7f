"""GPU tests of the K-loop's launch chain: programmatic dependent launch, device-side loss coefficients, the deferred
hyper-parameter gradient reduction and the persistent operator split must not change a single bit of the results.
(All through the C ABI; the CPU oracle anchors the numbers in test_gpu_parity.py -- here the variants are compared
with each other, which is the stricter statement: bit equality.)"""
import pytest
import torch

from helpers import rel_l2, random_problem
from oracle import dadmm_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _dev(t):
    return t.squeeze(-1).contiguous().to(DEV) if t.dim() == 4 else t.contiguous().to(DEV)


def _setup(P, n, m, B, K, seed, two_stage):
    from dadmm_b200 import functional as DF
    from dadmm_b200.graph import BatchGraph
    pr = random_problem(P, n, m, B, K, seed=seed, a_scale=0.1)
    hyp = O.hyp_table(pr["param"], torch.tensor([0.1, 0.99, 0.99, 0.99]), True)
    A = pr["A"].to(DEV)
    W = DF.atx(A, A)[0].contiguous()
    Atb = DF.atx(A, pr["b"].to(DEV)).squeeze(-1)
    fac = (A[0].contiguous(), A[0].transpose(1, 2).contiguous(), pr["b"].to(DEV).squeeze(-1).contiguous()) if two_stage else None
    if two_stage:
        from dadmm_b200 import _lib
        assert _lib.lib.dadmm_unfolded_uses_factor(0, 0, B, P, n, m) == 1, "shape does not take the two-stage route"
    graph = BatchGraph.from_graph_list(pr["graphs"], P, DEV)
    clamps = [DF.clamps_model1(k) for k in range(K)]
    state = tuple(_dev(pr[k]) for k in ("y0", "U0", "d0"))
    return DF, pr, hyp, W, Atb, fac, graph, clamps, state


def _run(DF, pr, hyp, W, Atb, fac, graph, clamps, state, algo="f16", coefs="device", weights=None):
    """One fwd + fused-loss bwd through Unfolded / MSELoss.  coefs: how the loss coefficients reach the reverse sweep."""
    K = hyp.shape[0]
    h = hyp.to(DEV).requires_grad_(True)
    handle = DF.FusedLossHandle()
    if coefs == "host":               # force the host-side coefficient list (dadmm_unfolded_bwd: loss_coef)
        offer = handle.offer
        handle.offer = lambda label, c, sentinel: offer(label, [float(v) for v in c.cpu().tolist()], sentinel)
    Y = DF.Unfolded.apply(h, W, W, Atb, *state, graph, clamps, algo, None, handle, fac, None if fac is None else fac[:2])
    losses = DF.MSELoss.apply(Y, pr["label"].to(DEV), None, handle)
    w = torch.zeros(K, device=DEV) if weights is None else torch.tensor(weights, device=DEV)
    if weights is None:
        w[-1] = 1.0
    (losses * w).sum().backward()
    assert handle.pending is None          # the reverse sweep consumed the (label, coef) side channel
    return Y.detach().clone(), h.grad.detach().clone(), losses.detach().clone()


@pytest.mark.parametrize("two_stage", [False, True])
def test_programmatic_dependent_launch_is_bit_identical(two_stage):
    """Same forward / backward with the chain launched programmatically (default) and the classic way, alternating, on a
    shape that takes the fused tensor-core path (several waves of level CTAs per contraction, so early-scheduled CTAs
    really do sit behind a running predecessor)."""
    from dadmm_b200 import _lib
    cfg = _setup(P=6, n=512, m=160, B=384, K=10, seed=5, two_stage=two_stage)
    ref = None
    prev = _lib.set_pdl(True)
    try:
        for rep in range(6):
            _lib.set_pdl(rep % 2 == 0)
            out = _run(*cfg)
            if ref is None:
                ref = out
            for a, b in zip(out, ref):
                assert torch.equal(a, b), f"rep {rep} (pdl={'on' if rep % 2 == 0 else 'off'}) differs"
    finally:
        _lib.set_pdl(prev)
    assert torch.isfinite(ref[0]).all() and torch.isfinite(ref[1]).all() and float(ref[1].abs().sum()) > 0


def test_pdl_on_the_fma_path_and_small_shapes():
    """The level kernels behind the exact-FMA contraction (classic launches in front of chain launches) and a shape
    below every tensor-core tile: bit-identical with and without programmatic launches."""
    from dadmm_b200 import _lib
    prev = _lib.set_pdl(True)
    try:
        for shape, algo in (((3, 40, 8, 5, 4), "simt"), ((4, 256, 64, 256, 6), "simt"), ((4, 256, 64, 256, 6), "tc")):
            cfg = _setup(*shape, seed=9, two_stage=False)
            _lib.set_pdl(True)
            on = _run(*cfg, algo=algo)
            _lib.set_pdl(False)
            off = _run(*cfg, algo=algo)
            for a, b in zip(on, off):
                assert torch.equal(a, b), (shape, algo)
    finally:
        _lib.set_pdl(prev)


@pytest.mark.parametrize("weights", [None, "mean", "mixed"])
def test_device_side_loss_coefficients_equal_host_side(weights):
    """d loss / d losses[k] handed to the reverse sweep as a device vector (no host sync) == the host list."""
    K = 8
    w = {None: None, "mean": [1.0 / K] * K, "mixed": [0.0, 0.3, 0.0, 0.0, 0.2, 0.0, 0.0, 0.5]}[weights]
    cfg = _setup(P=4, n=512, m=160, B=256, K=K, seed=13, two_stage=True)
    dev = _run(*cfg, coefs="device", weights=w)
    host = _run(*cfg, coefs="host", weights=w)
    for a, b in zip(dev, host):
        assert torch.equal(a, b)
    # and both equal the dense-gradient path (no fused loss term at all) to rounding
    DF, pr, hyp, W, Atb, fac, graph, clamps, state = cfg
    h = hyp.to(DEV).requires_grad_(True)
    Y = DF.Unfolded.apply(h, W, W, Atb, *state, graph, clamps, "f16", None, None, fac, fac[:2])
    losses = DF.MSELoss.apply(Y, pr["label"].to(DEV), None, None)
    wt = torch.tensor(w, device=DEV) if w is not None else torch.nn.functional.one_hot(torch.tensor(K - 1), K).float().to(DEV)
    (losses * wt).sum().backward()
    assert rel_l2(dev[1], h.grad) < 1e-5


def test_deferred_gradient_reduction_matches_fp64_sweep():
    """ghyp [K,P,4] from the single post-sweep reduction against the fp64 instantiation of the same sweep, every level
    and every column (a mis-addressed level row would be off by O(1))."""
    DF, pr, hyp, W, Atb, fac, graph, clamps, state = _setup(P=5, n=128, m=32, B=96, K=7, seed=17, two_stage=False)
    g32 = _run(DF, pr, hyp, W, Atb, None, graph, clamps, state, algo="simt")[1]
    W64, Atb64 = W.double(), Atb.double()
    s64 = tuple(t.double() for t in state)
    pr64 = dict(pr, label=pr["label"].double())
    g64 = _run(DF, pr64, hyp.double(), W64, Atb64, None, graph, clamps, s64, algo="simt")[1]
    assert g32.shape == (7, 5, 4)
    assert rel_l2(g32, g64) < 1e-4
    gmax = float(g64.abs().max())
    for k in range(7):
        for c in range(4):
            ref = g64[k, :, c]
            err = float((g32[k, :, c].double() - ref).abs().max())
            assert err <= 1e-3 * float(ref.abs().max()) + 1e-5 * gmax, (k, c, err)
    assert float(g64[6, :, 3].abs().max()) == 0.0          # eta of the last iteration only feeds U_K, which nothing reads


def test_operator_split_cache_reuse_and_invalidation():
    """Persistent operator split: second and later calls reuse it (bit-identical to splitting every call); an in-place
    change of the operator invalidates it."""
    from dadmm_b200 import functional as DF
    cfg = _setup(P=3, n=512, m=160, B=256, K=6, seed=23, two_stage=True)
    cache = DF._op_splits
    assert cache.enabled
    cache.entries.clear()
    first = _run(*cfg)
    assert len(cache.entries) == 1 and all(e[1] for e in cache.entries.values())      # filled by the forward, reused by the sweep
    second = _run(*cfg)
    cache.enabled = False
    try:
        plain = _run(*cfg)
    finally:
        cache.enabled = True
    for a, b, c in zip(first, second, plain):
        assert torch.equal(a, b) and torch.equal(a, c)
    # in-place update of the factor pair: new version counter -> new entry -> results follow the new operator
    DF_, pr, hyp, W, Atb, fac, graph, clamps, state = cfg
    fac[0].mul_(0.5)
    fac[1].mul_(0.5)
    changed = _run(*cfg)
    cache.enabled = False
    try:
        changed_plain = _run(*cfg)
    finally:
        cache.enabled = True
    for a, b in zip(changed, changed_plain):
        assert torch.equal(a, b)
    assert not torch.equal(changed[0], first[0])
    assert len(cache.entries) == 2


def _module_case(P=3, n=512, m=160, B=256, K=6, seed=31):
    import argparse
    import unfolded_DLASSO
    pr = random_problem(P, n, m, B, K, seed=seed, a_scale=0.1)
    args = argparse.Namespace(m=m, n=n, P=P, GHN_iter_num=K, DADMM_mode="diff", alpha_max=0.1, tau_max=0.99, rho_max=0.99,
                              eta_max=0.99, max_penalty_threshold=0.8, penalty_reduction_factor=0.95, batch_size=B, snr=4)
    model = unfolded_DLASSO.DLASSO_unfolded(pr["A"].to(DEV), args).to(DEV)
    with torch.no_grad():
        model.seq_hyp.param.copy_(pr["param"])
    return model, pr


def test_module_skips_the_unused_Atb_on_the_two_stage_route(capsys):
    """When the library forms the residual as A^T (A y - b), the module no longer computes A^T b: results equal the route
    that subtracts a precomputed Atb to fp32 rounding, gradients included; and a NaN batch still reaches the guarded
    path (which needs Atb and computes it on demand)."""
    import gnn_dlasso_utils
    model, pr = _module_case()
    W = model._operators(torch.device(DEV))[1]
    assert model._residual_from_factor(W, len(pr["b"]))
    outs = []
    for rhs in (True, False):
        model.two_stage_rhs = rhs
        model.zero_grad()
        torch.manual_seed(2)
        Y, _ = model(pr["b"].to(DEV), pr["graphs"])
        _, lf = gnn_dlasso_utils.compute_loss(Y, pr["label"].to(DEV))
        lf.backward()
        outs.append((Y.detach().clone(), model.seq_hyp.param.grad.clone()))
    assert rel_l2(outs[0][0], outs[1][0]) < 1e-5 and rel_l2(outs[0][1], outs[1][1]) < 1e-4
    assert not torch.equal(outs[0][0], outs[1][0])          # really two different evaluation routes
    model.two_stage_rhs = True
    b = pr["b"].clone()
    b[1, 2, 0, 0] = float("nan")
    Y, _ = model(b.to(DEV), pr["graphs"])
    assert "NaN/Inf in gradient at iteration 0" in capsys.readouterr().out
    assert torch.isfinite(Y).all()


def test_training_steps_through_the_modules_are_reproducible_and_sync_free_in_backward():
    """Two identical 3-step Adam runs through the drop-in module (table -> Unfolded -> compute_loss -> backward -> step)
    give bit-identical parameters; the backward of a step enqueues without a host read of the loss gradient (the device
    coefficient path is the one taken: the side channel carries a tensor)."""
    import gnn_dlasso_utils
    from dadmm_b200 import functional as DF
    seen = []
    offer = DF.FusedLossHandle.offer

    def spy(self, label, coefs, sentinel):
        seen.append(type(coefs))
        return offer(self, label, coefs, sentinel)
    DF.FusedLossHandle.offer = spy
    try:
        finals = []
        for rep in range(2):
            model, pr = _module_case(K=5, seed=37)
            opt = torch.optim.Adam(model.parameters(), lr=1e-2)
            b, label = pr["b"].to(DEV), pr["label"].to(DEV)
            for step in range(3):
                torch.manual_seed(100 + step)
                Y, _ = model(b, pr["graphs"])
                _, lf = gnn_dlasso_utils.compute_loss(Y, label, check_finite=False)
                opt.zero_grad(set_to_none=True)
                lf.backward()
                opt.step()
            finals.append(model.seq_hyp.param.detach().clone())
    finally:
        DF.FusedLossHandle.offer = offer
    assert torch.equal(finals[0], finals[1])
    assert seen and all(t is torch.Tensor for t in seen)
