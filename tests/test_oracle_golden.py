"""Pin the CPU oracle (oracle/dadmm_oracle.py) against the golden vectors minted from the
UNMODIFIED reference (oracle/make_golden.py).  CPU only."""
import math

import pytest
import torch

from helpers import MODEL1_CASES, MODEL1_EXTRA_CASES, MODEL3_CASES, Golden, rel_l2
from oracle import dadmm_oracle as O


def _hyp_for(g, dtype=torch.float32):
    param = g.t("param", dtype)
    hyp = O.hyp_table(param, g.t("max_param", dtype), bool(g.z["training"]))
    if hyp.shape[1] == 1:
        hyp = hyp.expand(-1, g.P, -1)
    return hyp


@pytest.mark.parametrize("name", MODEL1_CASES + MODEL1_EXTRA_CASES)
def test_model1_forward_bit_exact(name):
    """Loop-order oracle == reference forward, bit for bit, in fp32."""
    g = Golden(name)
    A = g.t("A")
    AtA, Atb = O.atx(A, A), O.atx(A, g.t("b"))
    Y = O.unfolded_forward(AtA, Atb, g.graphs, g.t("y0"), g.t("U0"), g.t("d0"), _hyp_for(g), exact_delta=True)
    assert Y.shape == g.t("Y").shape
    assert torch.equal(Y, g.t("Y")), f"max abs diff {float((Y - g.t('Y')).abs().max())}"
    lm, lf = O.loss(Y, g.t("label"))
    assert math.isclose(float(lm), float(g.z["loss_mean"]), rel_tol=1e-6)
    assert math.isclose(float(lf), float(g.z["loss_final"]), rel_tol=1e-6)
    assert torch.allclose(_hyp_for(g)[-1].unsqueeze(-1)[: g.t("hyp_last").shape[0]], g.t("hyp_last"), rtol=0, atol=0)


@pytest.mark.parametrize("name", MODEL1_CASES + MODEL1_EXTRA_CASES)
def test_model1_fp64_forward_and_grad(name):
    """Vectorised (dense 2L) oracle in fp64 vs the reference run in fp64: proves the algebra,
    including d loss_final / d param through all K iterations."""
    g = Golden(name)
    dt = torch.float64
    A = g.t("A", dt)
    AtA, Atb = O.atx(A, A), O.atx(A, g.t("b", dt))
    param = g.t("param", dt).clone().requires_grad_(True)
    hyp = O.hyp_table(param, g.t("max_param", dt), bool(g.z["training"]))
    if hyp.shape[1] == 1:
        hyp = hyp.expand(-1, g.P, -1)
    Y = O.unfolded_forward(AtA, Atb, g.graphs, g.t("y0", dt), g.t("U0", dt), g.t("d0", dt), hyp)
    assert rel_l2(Y[-1], g.t("Y64")[-1]) < 1e-9
    _, lf = O.loss(Y, g.t("label", dt))
    lf.backward()
    assert math.isclose(float(lf.detach()), float(g.z["loss_final64"]), rel_tol=1e-10)
    assert rel_l2(param.grad, g.t("dparam64")) < 1e-7


@pytest.mark.parametrize("name", MODEL1_CASES + MODEL1_EXTRA_CASES)
def test_model1_fp32_grad_noise_floor(name):
    """fp32 oracle gradient vs reference fp32 gradient: same order of error as the reference's
    own fp32-vs-fp64 gap (SURVEY.md 8c protocol)."""
    g = Golden(name)
    A = g.t("A")
    AtA, Atb = O.atx(A, A), O.atx(A, g.t("b"))
    param = g.t("param").clone().requires_grad_(True)
    hyp = O.hyp_table(param, g.t("max_param"), bool(g.z["training"]))
    if hyp.shape[1] == 1:
        hyp = hyp.expand(-1, g.P, -1)
    Y = O.unfolded_forward(AtA, Atb, g.graphs, g.t("y0"), g.t("U0"), g.t("d0"), hyp)
    _, lf = O.loss(Y, g.t("label"))
    lf.backward()
    ref_gap = rel_l2(g.t("dparam"), g.t("dparam64"))
    ours = rel_l2(param.grad, g.t("dparam64"))
    assert ours <= max(1e-5, 4 * ref_gap), (ours, ref_gap)


@pytest.mark.parametrize("name", MODEL3_CASES)
def test_model3_recurrence_frozen_hyp(name):
    """Model #3 recurrence (fixed clamps, delta clamp) with the hypernetwork outputs frozen."""
    g = Golden(name)
    A = g.t("A")
    AtA, Atb = O.atx(A, A), O.atx(A, g.t("b"))
    Y = O.unfolded_forward(AtA, Atb, g.graphs, g.t("y0"), g.t("U0"), g.t("d0"), g.t("hyp"),
                           clamp_fn=O.clamps_model3, exact_delta=True)
    assert torch.equal(Y, g.t("Y")), f"max abs diff {float((Y - g.t('Y')).abs().max())}"


def test_delta_is_twice_laplacian():
    g = Golden("m1_same_pergraph_P8_n48")
    y = g.t("y0")
    d1 = O.delta_events(g.graphs, y)
    d2 = O.delta_dense(O.laplacian2(g.graphs, g.P), y)
    assert rel_l2(d1, d2) < 1e-6
    deg = O.degrees(g.graphs, g.P)
    assert torch.equal(O.laplacian2(g.graphs, g.P).diagonal(dim1=1, dim2=2), 2 * deg[:, :, 0, 0])


@pytest.mark.ref
def test_oracle_equals_reference_on_random_configurations():
    """Beyond the committed fixtures: 24 seeded random configurations (agents, sizes, K, 'same'/'diff' tables, train /
    eval mode, shared / per-problem graphs with and without bridging, conditioning of A) run through the UNMODIFIED
    reference class here and through the oracle -- forward bit for bit in fp32, loss to 1e-6, d loss / d param in fp64
    to 1e-7.  (Skipped where the reference checkout is absent.)"""
    import argparse
    import random
    import networkx as nx
    from oracle import ref_harness
    ref = ref_harness.load("unfolded_DLASSO")
    utils = ref_harness.load("gnn_dlasso_utils")
    rnd = random.Random(2024)
    for case in range(24):
        P, n = rnd.randint(2, 9), rnd.randint(3, 40)
        m, K, B = rnd.randint(1, max(1, n // 2)), rnd.randint(1, 9), rnd.randint(1, 4)
        mode, training = rnd.choice(["diff", "same"]), rnd.choice([True, False])
        args = argparse.Namespace(m=m, n=n, P=P, GHN_iter_num=K, DADMM_mode=mode, alpha_max=rnd.choice([0.05, 0.1]), tau_max=0.99,
                                  rho_max=0.99, eta_max=0.99, max_penalty_threshold=0.8, penalty_reduction_factor=0.95,
                                  batch_size=B, snr=4)
        torch.manual_seed(case)
        A = utils.set_A(args) * rnd.choice([1.0, 0.1])
        label = 2 * torch.randn(B, n, 1) * (torch.rand(B, n, 1) <= 0.25)
        b = torch.stack([A[0, p] @ label for p in range(P)], dim=1)
        if rnd.random() < 0.4:
            graphs = [nx.erdos_renyi_graph(P, 0.5, seed=case)] * B
        else:
            graphs = []
            for i in range(B):
                g = nx.erdos_renyi_graph(P, rnd.choice([0.15, 0.5, 0.9]), seed=100 * case + i)
                if rnd.random() < 0.5 and not nx.is_connected(g):
                    comps = list(nx.connected_components(g))
                    for c in range(len(comps) - 1):
                        g.add_edge(list(comps[c])[0], list(comps[c + 1])[0])
                graphs.append(g)
        param = torch.randn(K, 1 if mode == "same" else P, 4) * 0.5 + 0.3
        mp = torch.tensor([args.alpha_max, 0.99, 0.99, 0.99])
        outs = {}
        for dt in (torch.float32, torch.float64):
            with ref_harness.default_dtype(dt):
                model = ref.DLASSO_unfolded(A.to(dt), args)
                model.train(training)
                with torch.no_grad():
                    model.seq_hyp.param.copy_(param.to(dt))
                torch.manual_seed(1000 + case)
                Y, _ = model(b.to(dt), graphs)
                lm, lf = utils.compute_loss(Y, label.to(dt))
                lf.backward()
                outs[dt] = (Y.detach(), float(lm.detach()), float(lf.detach()), model.seq_hyp.param.grad.detach().clone())
        torch.manual_seed(1000 + case)
        noise = [torch.randn((B, P, n, 1)) for _ in range(3)]
        # fp32, the reference's accumulation order: bit for bit
        hyp = O.hyp_table(param, mp, training)
        hyp = hyp.expand(-1, P, -1) if hyp.shape[1] == 1 else hyp
        Yo = O.unfolded_forward(O.atx(A, A), O.atx(A, b), graphs, *(t * 1e-2 for t in noise), hyp, exact_delta=True)
        assert torch.equal(Yo, outs[torch.float32][0]), (case, P, n, m, K, B, mode, training)
        lm, lf = O.loss(Yo, label)
        assert math.isclose(float(lm), outs[torch.float32][1], rel_tol=1e-6) and math.isclose(float(lf), outs[torch.float32][2], rel_tol=1e-6)
        # fp64, differentiable form: algebra incl. the gradient through all K iterations
        dt = torch.float64
        with ref_harness.default_dtype(dt):  # the reference's own fp64 draw (a different stream from fp32's)
            torch.manual_seed(1000 + case)
            noise64 = [torch.randn((B, P, n, 1)) for _ in range(3)]
        mp64 = torch.tensor([args.alpha_max, 0.99, 0.99, 0.99], dtype=dt)
        p64 = param.to(dt).clone().requires_grad_(True)
        hyp = O.hyp_table(p64, mp64, training)
        hyp = hyp.expand(-1, P, -1) if hyp.shape[1] == 1 else hyp
        A64 = A.to(dt)
        Y64 = O.unfolded_forward(O.atx(A64, A64), O.atx(A64, b.to(dt)), graphs, *(t * 1e-2 for t in noise64), hyp)
        _, lf64 = O.loss(Y64, label.to(dt))
        lf64.backward()
        assert rel_l2(Y64.detach(), outs[dt][0]) < 1e-9, case
        ref_g = outs[dt][3]
        if float(ref_g.abs().max()) > 0:
            assert rel_l2(p64.grad, ref_g) < 1e-7, (case, rel_l2(p64.grad, ref_g))


@pytest.mark.ref
def test_committed_fixtures_are_what_the_committed_script_mints(tmp_path, monkeypatch, capsys):
    """``oracle/make_golden.py`` run again against the unmodified reference reproduces every array of every committed
    fixture bit for bit (inputs, fp32 and fp64 outputs, gradients).  (Skipped where the reference checkout is absent.)"""
    import os
    import numpy as np
    from helpers import ROOT
    from oracle import make_golden
    monkeypatch.setattr(make_golden, "OUT", str(tmp_path))
    monkeypatch.setattr("sys.argv", ["make_golden.py"])
    state = torch.random.get_rng_state()
    try:
        make_golden.main()
    finally:
        torch.random.set_rng_state(state)
    capsys.readouterr()
    committed = os.path.join(ROOT, "tests", "golden")
    names = sorted(f for f in os.listdir(committed) if f.endswith(".npz"))
    assert names == sorted(os.listdir(tmp_path)) and len(names) >= 6
    for f in names:
        a, b = np.load(os.path.join(committed, f), allow_pickle=True), np.load(os.path.join(tmp_path, f), allow_pickle=True)
        assert set(a.files) == set(b.files), f
        for k in a.files:
            x, y = a[k], b[k]
            assert x.shape == y.shape and x.dtype == y.dtype, (f, k)
            if x.dtype.kind in "fiub":
                assert np.array_equal(x, y, equal_nan=True), (f, k)
            else:
                assert x.tolist() == y.tolist(), (f, k)
