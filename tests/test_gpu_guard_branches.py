"""The four NaN/Inf guard branches of ``DLASSO_unfolded.forward`` (unfolded_DLASSO.py:55-61 y_k / U_k reset, :84-86 gradient
skipped, :102-104 y_next replaced by the previous iterate) against the UNMODIFIED reference class run on the CPU on the
same inputs (the reference travels to the GPU box as ``oracle/_ref``, see oracle/build_ref.py).  Each case plants one
non-finite value where only that branch can catch it; the drop-in module must print the same warnings and return the
same iterates -- including which entries are left non-finite, if any."""
import argparse

import pytest
import torch

from helpers import random_problem, rel_l2
from oracle import ref_harness as RH

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
P, n, m, B, K = 5, 64, 16, 6, 5


def _args():
    return argparse.Namespace(m=m, n=n, P=P, GHN_iter_num=K, DADMM_mode="diff", alpha_max=0.1, tau_max=0.99, rho_max=0.99,
                              eta_max=0.99, max_penalty_threshold=0.8, penalty_reduction_factor=0.95, batch_size=B, snr=4)


def _warnings(text):
    return sorted({ln.strip() for ln in text.splitlines() if ln.startswith("Warning")})


@pytest.mark.parametrize("case", ["y0_nan", "U0_inf", "b_nan", "alpha_nan_iteration_2", "alpha_nan_last_iteration", "eta_nan_iteration_1"])
def test_guard_branch_matches_the_reference(case, capsys):
    if not RH.reference_root("unfolded_DLASSO"):
        pytest.skip("reference modules neither at /root/reference nor staged under oracle/_ref")
    import unfolded_DLASSO
    ref_mod = RH.load("unfolded_DLASSO")
    pr = random_problem(P, n, m, B, K, seed=23, a_scale=0.1)
    gen = torch.Generator().manual_seed(5)
    noise = [torch.randn((B, P, n, 1), generator=gen) for _ in range(3)]
    b, param = pr["b"].clone(), pr["param"].clone()
    if case == "y0_nan":
        noise[0][2, 1, 7, 0] = float("nan")
    elif case == "U0_inf":
        noise[1][0, 3, 11, 0] = float("inf")
    elif case == "b_nan":
        b[1, 2, 0, 0] = float("nan")
    elif case == "alpha_nan_iteration_2":
        param[2, 1, 0] = float("nan")
    elif case == "alpha_nan_last_iteration":
        param[K - 1, 0, 0] = float("nan")
    elif case == "eta_nan_iteration_1":
        param[1, 4, 3] = float("nan")
    # reference on the CPU, its three randn draws replaced by the same numbers
    ref = ref_mod.DLASSO_unfolded(pr["A"], _args())
    with torch.no_grad():
        ref.seq_hyp.param.copy_(param)
    it, orig = iter(noise), torch.randn
    torch.randn = lambda *a, **k: next(it).clone()
    try:
        with torch.no_grad():
            Yr, _ = ref(b, pr["graphs"])
    finally:
        torch.randn = orig
    w_ref = _warnings(capsys.readouterr().out)
    model = unfolded_DLASSO.DLASSO_unfolded(pr["A"].to(DEV), _args()).to(DEV)
    with torch.no_grad():
        model.seq_hyp.param.copy_(param.to(DEV))
        Y, _ = model(b.to(DEV), pr["graphs"], noise=[t.to(DEV) * 1e-2 for t in noise])
    w_ours = _warnings(capsys.readouterr().out)
    Y = Y.cpu()
    assert w_ref, "the planted value did not reach a guard in the reference"
    assert w_ours == w_ref, (w_ours, w_ref)
    assert torch.equal(torch.isfinite(Y), torch.isfinite(Yr))
    fin = torch.isfinite(Yr)
    assert rel_l2(Y[fin], Yr[fin]) < 1e-5
