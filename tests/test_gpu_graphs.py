"""Whole-step CUDA-graph capture (``dadmm_b200.graphs.GraphedStep``): a captured training step -- forward through the fused
K-loop, ``compute_loss``, backward, Adam -- replays bit-identically to the same step launched eagerly, on the host-bound
shape of BASELINE configs[0] and on a shape that takes the tensor-core route; the sticky non-finite flags of the
``check_finite = "deferred"`` mode report a planted NaN after the replay."""
import argparse

import pytest
import torch

from helpers import random_problem

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _case(P, n, m, B, K, seed):
    import unfolded_DLASSO
    pr = random_problem(P, n, m, B, K, seed=seed, a_scale=0.1, per_sample_graphs=(P != 5))
    args = argparse.Namespace(m=m, n=n, P=P, GHN_iter_num=K, DADMM_mode="diff", alpha_max=0.1, tau_max=0.99, rho_max=0.99,
                              eta_max=0.99, max_penalty_threshold=0.8, penalty_reduction_factor=0.95, batch_size=B, snr=4)
    model = unfolded_DLASSO.DLASSO_unfolded(pr["A"].to(DEV), args).to(DEV)
    with torch.no_grad():
        model.seq_hyp.param.copy_(pr["param"])
    return model, pr


@pytest.mark.parametrize("shape", [(5, 500, 100, 32, 15), (3, 512, 160, 256, 5)])
def test_graphed_training_step_is_bit_identical_to_eager(shape):
    import gnn_dlasso_utils
    from dadmm_b200.graphs import GraphedStep
    P, n, m, B, K = shape
    model, pr = _case(P, n, m, B, K, seed=17)
    model.check_finite = "deferred"
    optim = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True)
    graphs = pr["graphs"]
    inputs = [pr["b"].to(DEV), pr["label"].to(DEV)] + [pr[k].to(DEV) for k in ("y0", "U0", "d0")]
    start = model.seq_hyp.param.detach().clone()

    def step(b, label, y0, U0, d0):
        Y, _ = model(b, graphs, noise=(y0, U0, d0))
        _, lf = gnn_dlasso_utils.compute_loss(Y, label, check_finite=False)
        optim.zero_grad(set_to_none=True)
        lf.backward()
        optim.step()
        return Y, lf.detach(), model.seq_hyp.param.grad

    def run(fn, steps):
        with torch.no_grad():
            model.seq_hyp.param.copy_(start)
            for st in optim.state.values():           # in place: the captured graph updates these very tensors
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()
        out = []
        for _ in range(steps):
            Y, lf, g = fn(*inputs)
            out.append((Y.detach().clone(), lf.clone(), g.clone(), model.seq_hyp.param.detach().clone()))
        return out

    eager = run(step, 3)
    graphed = GraphedStep(step, inputs, warmup=2)
    # run() resets the parameter and the optimizer state the warm-up and the capture advanced, in place
    replay = run(graphed, 3)
    for i, (a, b) in enumerate(zip(eager, replay)):
        for name, x, y in zip(("Y", "loss", "grad", "param"), a, b):
            assert torch.equal(x, y), (i, name, float((x - y).abs().max()))
    assert not model.nonfinite_seen()
    bad = [t.clone() for t in inputs]
    bad[0][0, 0, 0, 0] = float("nan")
    graphed(*bad)
    assert model.nonfinite_seen() and not model.nonfinite_seen()      # reported once, then cleared
