"""Drop-in replacement of the reference ``gnn_dlasso_utils.py``: ``set_A``, ``compute_loss``, ``compute_loss2``.

``compute_loss`` (reference gnn_dlasso_utils.py:27-88) is the per-iteration, per-agent MSE; the K*P
``F.mse_loss`` slices (and the O(K*P) full-size zero tensors autograd allocates behind them) are
replaced by one reduction kernel and a backward that is fused into the solver's reverse sweep
(``dadmm_b200.functional.MSELoss``).  Semantics kept: returns ``(mean_k losses + 1e-8,
losses[K-1] + 1e-8)``; non-finite inputs / losses return the constant 1.0 pair with a printed warning.
"""
import torch

from dadmm_b200 import functional as DF


def set_A(args):
    """Problem matrices A [1,P,m,n] on the CPU (reference :4-16): per agent a Gaussian matrix whose
    singular values are clipped to [0.1, 10].  Same RNG consumption (one randn(m,n) per agent)."""
    A = torch.zeros((1, args.P, args.m, args.n))
    for agent in range(args.P):
        left, sv, right = torch.svd(torch.randn((args.m, args.n)))
        A[0, agent] = left @ torch.diag(sv.clamp(min=0.1, max=10.0)) @ right.T
    return A


def compute_loss2(Y, label):
    """|label|-weighted MSE of the agent-averaged iterates (reference :18-25); off the hot path."""
    w = label.abs() + 0.0001
    w = w / w.sum(dim=1).unsqueeze(-1)
    y_mean = Y.mean(dim=2)
    per_elem = lambda est: ((est - label) ** 2 * w).sum(dim=1)
    return per_elem(y_mean.mean(dim=0)).mean(), per_elem(y_mean[-1]).mean()


def _one(device):
    return torch.tensor(1.0, device=device), torch.tensor(1.0, device=device)


def compute_loss(Y, label, check_finite=True, global_batch=None):
    """Y [K,B,P,n,1], label [B,n,1] -> (loss_mean, loss_final).

    ``global_batch``: total batch size when the batch is sharded over ranks (the per-rank losses then
    sum to the global loss).

    NaN guards (reference :29-47, :68-82: non-finite Y / label / losses return the constant 1.0 pair with a printed
    warning).  The reference pays three full-tensor scans and three host syncs for them; here
      * Y needs no scan when it comes straight from the solver modules: their forward has already read the kernels'
        non-finite flags (``Y._dadmm_finite`` records the tensor version that check covered -- any later in-place
        write to Y voids it and the scan below runs);
      * label ([B,n]) and the K losses are checked with ONE host read; only a failure looks closer.
    ``check_finite=False`` skips the guards altogether."""
    DF.require_cuda(Y, label)
    if check_finite and getattr(Y, "_dadmm_finite", None) != Y._version:
        if not bool(torch.isfinite(Y).all()):
            print("Warning: NaN/Inf detected in model output Y")
            return _one(Y.device)
    losses = DF.MSELoss.apply(Y, label.to(Y.dtype), global_batch, getattr(Y, "_dadmm_handle", None))
    if check_finite and not bool(torch.isfinite(label).all() & torch.isfinite(losses).all()):
        if not bool(torch.isfinite(label).all()):
            print("Warning: NaN/Inf detected in label")
        else:
            print("Warning: NaN/Inf detected in computed losses")
        return _one(Y.device)
    eps = 1e-8
    return losses.mean() + eps, losses[-1] + eps
