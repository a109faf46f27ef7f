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
