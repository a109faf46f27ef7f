"""GPU tests of the NaN guards around the solver: ``compute_loss`` with the reference's guards on (the drivers' call)
reuses the non-finite flags the modules' forward has already read from the kernels."""
import pytest
import torch

from test_gpu_chain import DEV, _module_case

pytestmark = pytest.mark.gpu


def test_compute_loss_guards_reuse_the_solver_flags(capsys):
    """compute_loss with the reference's NaN guards on (the drivers' call): Y straight from the module is not scanned
    again (its forward already read the kernels' non-finite flags), label and losses cost one small host read; an
    in-place write to Y voids the shortcut and the full scan catches a NaN planted afterwards."""
    import gnn_dlasso_utils
    model, pr = _module_case(K=4, seed=41)
    b, label = pr["b"].to(DEV), pr["label"].to(DEV)
    Y, _ = model(b, pr["graphs"])
    assert Y._dadmm_finite == Y._version
    seen, orig = [], torch.isfinite

    def spy(t):
        seen.append(t.numel())
        return orig(t)
    torch.isfinite = spy
    try:
        lm, lf = gnn_dlasso_utils.compute_loss(Y, label)
    finally:
        torch.isfinite = orig
    assert seen and max(seen) <= label.numel()
    lm0, lf0 = gnn_dlasso_utils.compute_loss(Y, label, check_finite=False)
    assert float(lm) == float(lm0) and float(lf) == float(lf0)
    lf.backward()
    assert torch.isfinite(model.seq_hyp.param.grad).all()
    Y2 = Y.detach().clone()
    Y2._dadmm_finite = Y2._version
    Y2[0, 0, 0, 0, 0] = float("nan")
    one = gnn_dlasso_utils.compute_loss(Y2, label)
    assert "NaN/Inf detected in model output Y" in capsys.readouterr().out
    assert float(one[0]) == 1.0 and float(one[1]) == 1.0
    bad = label.clone()
    bad[0, 0, 0] = float("inf")
    one = gnn_dlasso_utils.compute_loss(Y, bad)
    assert "NaN/Inf detected in label" in capsys.readouterr().out and float(one[1]) == 1.0
