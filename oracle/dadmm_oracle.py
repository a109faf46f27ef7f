"""CPU oracle for the unfolded D-ADMM distributed-LASSO hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline.  The product
path (``hyperparameter-gnn_unfolded-d-admm-main_b200/``) never imports anything
from ``oracle/`` and raises when the CUDA library is missing.

It restates, on the CPU (torch CPU tensors, fp32 or fp64), the algorithm of the
reference's two unfolded solvers; every function cites the reference file:line it
follows (paths relative to the reference checkout):

* ``unfolded_DLASSO.py:34-110``  (``DLASSO_unfolded.forward``)          -> ``unfolded_forward``
* ``unfolded_DLASSO.py:156-168`` (``seq_hyperparam.forward``)           -> ``hyp_table``
* ``unfolded_DLASSO.py:111-118`` (``compute_sum_neighbors``)            -> ``degrees``
* ``unfolded_DLASSO.py:120-124`` (``compute_Atx``)                      -> ``atx``
* ``unfolded_DLASSO.py:127-140`` (``compute_delta``)                    -> ``delta_events`` (bit-faithful
  accumulation order) and ``delta_dense`` (``2*L @ y``, vectorised)
* ``gnn_dlasso_models_progressive.py:198-232`` (model #3 recurrence)    -> ``step`` with fixed clamps
* ``gnn_dlasso_utils.py:27-88``  (``compute_loss``)                     -> ``loss``
* ``gnn_dlasso_utils.py:4-16``   (``set_A``)                            -> ``set_A``

Pinning: ``oracle/make_golden.py`` runs the UNMODIFIED reference classes (imported
from ``/root/reference`` in the build container, ``torch_geometric`` stubbed) and
stores their outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks
this file against those vectors (bit-exact forward for the loop form in fp32).
The GNN hypernetwork of model #3 depends on ``torch_geometric`` (un-vendored,
unpinned: reference ``requirements.txt:11``) -- for that part **parity is unpinned**;
the recurrence inside model #3 is pinned with a frozen per-sample hyper-parameter
tensor.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

INF = float("inf")


# --------------------------------------------------------------------------------------
# problem generator
# --------------------------------------------------------------------------------------
def set_A(args) -> torch.Tensor:
    """gnn_dlasso_utils.py:4-16: per agent ``randn(m,n)`` -> SVD -> singular values clamped to [0.1, 10] ->
    ``U diag(S) V^T``; A [1,P,m,n] fp32 on the CPU.  Same RNG consumption as the reference (one draw per agent)."""
    A = torch.zeros((1, args.P, args.m, args.n))
    for p in range(args.P):
        U, S, V = torch.svd(torch.randn((args.m, args.n)))
        A[0, p] = U @ torch.diag(torch.clamp(S, min=0.1, max=10.0)) @ V.T
    return A


# --------------------------------------------------------------------------------------
# graph helpers
# --------------------------------------------------------------------------------------
def neighbor_lists(graph, P: int) -> List[List[int]]:
    """Adjacency lists in the iteration order of ``graph.neighbors(p)`` (what the
    reference loops over, unfolded_DLASSO.py:136)."""
    return [list(graph.neighbors(p)) for p in range(P)]


def event_lists(graph, P: int) -> List[List[int]]:
    """Per node q: the neighbour ids in the exact order in which the reference's
    ``compute_delta`` (unfolded_DLASSO.py:132-139) adds ``y_q - y_e`` into ``delta[b,q]``.

    The reference walks p = 0..P-1 and, for j in N(p): ``delta[p] += y_p - y_j`` and
    ``delta[j] -= y_p - y_j``.  ``x - (a-b) == x + (b-a)`` bit-for-bit in IEEE arithmetic, so
    node q receives, in this order: one term per neighbour p<q (while the outer loop
    is at p), then one term per j in N(q) in adjacency order (outer loop at q), then
    one term per neighbour p>q.  Every undirected edge therefore contributes twice
    (delta = 2*L*y)."""
    nbrs = neighbor_lists(graph, P)
    ev: List[List[int]] = [[] for _ in range(P)]
    for p in range(P):
        for j in nbrs[p]:
            ev[p].append(j)          # delta[p] += y_p - y_j
            ev[j].append(p)          # delta[j] -= y_p - y_j  ==  delta[j] += y_j - y_p
    return ev


def degrees(graph_list: Sequence, P: int, dtype=torch.float32) -> torch.Tensor:
    """unfolded_DLASSO.py:111-118: ``deg[b,p] = len(list(graph.neighbors(p)))`` -> [B,P,1,1]."""
    cache = {}
    out = torch.zeros((len(graph_list), P, 1, 1), dtype=dtype)
    for i, g in enumerate(graph_list):
        key = id(g)
        if key not in cache:
            cache[key] = torch.tensor([len(list(g.neighbors(p))) for p in range(P)], dtype=dtype)
        out[i, :, 0, 0] = cache[key]
    return out


def laplacian2(graph_list: Sequence, P: int, dtype=torch.float32) -> torch.Tensor:
    """Dense ``2*L`` per sample, [B,P,P], such that ``delta = 2L @ y`` reproduces
    unfolded_DLASSO.py:127-140 up to summation order."""
    cache = {}
    out = torch.zeros((len(graph_list), P, P), dtype=dtype)
    for i, g in enumerate(graph_list):
        key = id(g)
        if key not in cache:
            M = torch.zeros((P, P), dtype=dtype)
            for q, ev in enumerate(event_lists(g, P)):
                for e in ev:
                    M[q, q] += 1.0
                    M[q, e] -= 1.0
            cache[key] = M
        out[i] = cache[key]
    return out


def delta_events(graph_list: Sequence, y: torch.Tensor) -> torch.Tensor:
    """Bit-faithful restatement of ``compute_delta`` (unfolded_DLASSO.py:127-140): sequential
    accumulation in the reference's order.  y: [B,P,n,1] -> delta [B,P,n,1].  No autograd."""
    B, P = y.shape[0], y.shape[1]
    out = torch.zeros_like(y)
    cache = {}
    for b in range(B):
        g = graph_list[b]
        if id(g) not in cache:
            cache[id(g)] = event_lists(g, P)
        ev = cache[id(g)]
        for q in range(P):
            acc = torch.zeros_like(y[b, q])
            for e in ev[q]:
                acc = acc + (y[b, q] - y[b, e])
            out[b, q] = acc
    return out


def delta_dense(lap2: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """``delta = 2L y`` vectorised (differentiable).  lap2 [B,P,P], y [B,P,n,1]."""
    return torch.einsum("bpq,bqnk->bpnk", lap2, y)


# --------------------------------------------------------------------------------------
# contractions
# --------------------------------------------------------------------------------------
def atx(A: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """unfolded_DLASSO.py:120-124: ``Atx[:,p] = A[0,p].T @ x[:,p]``; A [1,P,m,n], x [Bx,P,m,c]."""
    P = A.shape[1]
    out = torch.zeros((x.shape[0], P, A.shape[3], x.shape[3]), dtype=x.dtype)
    for p in range(P):
        out[:, p] = torch.matmul(A[0, p].T, x[:, p])
    return out


def contract(AtA: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """unfolded_DLASSO.py:69-71: ``AtAy[:,p] = AtA[0,p] @ y[:,p]``; AtA [1,P,n,n], y [B,P,n,1]."""
    out = torch.zeros_like(y)
    for p in range(AtA.shape[1]):
        out[:, p] = torch.matmul(AtA[0, p], y[:, p])
    return out


def contract_gemm(AtA: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Same product as ``contract`` as ONE batched GEMM per call (``AtA_p [n,n] @ Y_p [n,B]``) instead of the reference's
    B broadcast mat-vecs per agent -- identical up to summation order (1e-15 in fp64, where this form is used: the fp64
    checker of the GPU tests at the BASELINE shapes, which would otherwise spend minutes in memory-bound GEMVs)."""
    return torch.einsum("pij,bpj->bpi", AtA[0], y[..., 0]).unsqueeze(-1)


# --------------------------------------------------------------------------------------
# hyper-parameter table (model #1)
# --------------------------------------------------------------------------------------
def hyp_table(param: torch.Tensor, max_param: torch.Tensor, training: bool,
              max_penalty_threshold: float = 0.8, penalty_reduction_factor: float = 0.95) -> torch.Tensor:
    """unfolded_DLASSO.py:156-168 for every k at once -> [K, P|1, 4] (columns alpha,tau,rho,eta).

    h_k = sigmoid(sum_{i<=k} param[i]) * max_param; in training mode, if mean(h_k) > threshold:
    h_k *= factor; clamp to [1e-4, 0.99]."""
    K = param.shape[0]
    rows = []
    for k in range(K):
        h = torch.sum(param[:k + 1], dim=0)                  # [P|1, 4]
        h = torch.sigmoid(h) * max_param.reshape(1, 4).to(h.dtype)
        if training:
            pen = torch.sum(h) / (h.shape[0] * h.shape[1])
            if pen > max_penalty_threshold:
                h = h * penalty_reduction_factor
        rows.append(torch.clamp(h, min=1e-4, max=0.99))
    return torch.stack(rows)


# --------------------------------------------------------------------------------------
# clamp schedules
# --------------------------------------------------------------------------------------
@dataclass
class Clamps:
    """Element-wise clamp bounds of one iteration: grad (G), y (V), delta (D), U (Uc)."""
    G: float
    V: float
    D: float
    Uc: float


def clamps_model1(k: int) -> Clamps:
    """unfolded_DLASSO.py:80 ``max(1, 30-k)``, :92 ``max(10, 200-3k)``; no delta clamp (:96 commented)."""
    v = max(10.0, 200.0 - k * 3)
    return Clamps(G=max(1.0, 30.0 - k), V=v, D=INF, Uc=v)


def clamps_model3(k: int = 0) -> Clamps:
    """gnn_dlasso_models_progressive.py:212,224,229,232: fixed 10 / 100 / 20 / 100."""
    return Clamps(G=10.0, V=100.0, D=20.0, Uc=100.0)


# --------------------------------------------------------------------------------------
# one iteration and the K-iteration recurrence
# --------------------------------------------------------------------------------------
def step(AtAy, Atb, deg, y, U, delta, alpha, tau, rho, eta, c: Clamps, delta_fn):
    """One D-ADMM iteration, unfolded_DLASSO.py:73-99 (model #3: ...progressive.py:205-232).

    All state tensors [B,P,n,1]; alpha..eta broadcastable ([1,P,1,1] or [B,P,1,1]).
    Returns (y_next, U_next, delta_next, grad_raw).  The 5-term sum is evaluated left to
    right exactly as written in the reference."""
    grad_raw = AtAy - Atb + y.sign() * tau + U * deg + delta * rho
    grad = torch.clamp(grad_raw, -c.G, c.G)
    y_next = torch.clamp(y - alpha * grad, -c.V, c.V)
    d_next = delta_fn(y_next)
    if math.isfinite(c.D):
        d_next = torch.clamp(d_next, -c.D, c.D)
    U_next = torch.clamp(U + d_next * eta, -c.Uc, c.Uc)
    return y_next, U_next, d_next, grad_raw


def unfolded_forward(AtA, Atb, graph_list, y0, U0, d0, hyp, clamp_fn=clamps_model1,
                     exact_delta: bool = False, keep: bool = False, gemm_contract: bool = False):
    """K-iteration recurrence of ``DLASSO_unfolded.forward`` (unfolded_DLASSO.py:53-109) given
    the initial noise (y0,U0,d0) and a hyper-parameter table ``hyp`` [K,P,4] (shared over the
    batch, model #1) or [K,B,P,4] (per sample, model #3 with frozen hypernetwork output).

    exact_delta=True uses the reference's accumulation order (bit-faithful, no autograd);
    otherwise the dense 2L einsum (differentiable).  Returns Y [K,B,P,n,1] (+ per-iteration
    (y,U,delta,grad_raw,AtAy) inputs when keep=True).  gemm_contract=True evaluates the contraction as one batched GEMM
    (``contract_gemm``: same values up to summation order).  The NaN/Inf guards of
    unfolded_DLASSO.py:55-61,84-86,102-104 are no-ops for finite data and are omitted here;
    ``tests/test_nan_guard.py`` covers them through the guarded path."""
    B, P = y0.shape[0], y0.shape[1]
    dtype = y0.dtype
    deg = degrees(graph_list, P, dtype)
    if exact_delta:
        dfn = lambda y: delta_events(graph_list, y)
    else:
        lap2 = laplacian2(graph_list, P, dtype)
        dfn = lambda y: delta_dense(lap2, y)
    y, U, d = y0, U0, d0
    Y, trace = [], []
    K = hyp.shape[0]
    for k in range(K):
        h = hyp[k]
        if h.dim() == 2:      # [P,4] -> [1,P,1,1]
            al, ta, rh, et = (h[:, i].reshape(1, P, 1, 1) for i in range(4))
        else:                 # [B,P,4] -> [B,P,1,1]
            al, ta, rh, et = (h[:, :, i].reshape(B, P, 1, 1) for i in range(4))
        a = contract_gemm(AtA, y) if gemm_contract else contract(AtA, y)
        y_n, U_n, d_n, g_raw = step(a, Atb, deg, y, U, d, al, ta, rh, et, clamp_fn(k), dfn)
        if keep:
            trace.append(dict(y=y, U=U, delta=d, AtAy=a, grad_raw=g_raw, y_next=y_n, U_next=U_n, delta_next=d_n))
        y, U, d = y_n, U_n, d_n
        Y.append(y)
    Y = torch.stack(Y)
    return (Y, trace) if keep else Y


def loss(Y: torch.Tensor, label: torch.Tensor, vectorised: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """gnn_dlasso_utils.py:27-88 for finite inputs: ``losses[k] = mean_p mse(Y[k,:,p], label)``;
    returns (mean_k + 1e-8, last + 1e-8).  vectorised=True evaluates the K*P ``mse_loss`` slices as one reduction (same
    values up to summation order): autograd through the slice form allocates a full-size zero tensor per slice, which at
    the BASELINE shapes (K*P = 1250 slices of a GB-sized Y) is minutes of memory traffic in the fp64 checker."""
    K, B, P, n, _ = Y.shape
    Yr = Y.reshape(K, B, P, n)
    lab = label.reshape(B, n)
    if vectorised:
        losses = ((Yr - lab.reshape(1, B, 1, n)) ** 2).mean(dim=(1, 3)).sum(dim=1) / P
        return losses.mean() + 1e-8, losses[-1] + 1e-8
    losses = []
    for k in range(K):
        acc = 0.0
        for p in range(P):
            acc = acc + torch.mean((Yr[k, :, p] - lab) ** 2)
        losses.append(acc / P)
    losses = torch.stack(losses)
    return losses.mean() + 1e-8, losses[-1] + 1e-8


# --------------------------------------------------------------------------------------
# loop-faithful port used as the timed CPU baseline ("port" of the reference's cost model)
# --------------------------------------------------------------------------------------
def delta_loops_autograd(graph_list: Sequence, y: torch.Tensor) -> torch.Tensor:
    """Same Python triple loop + in-place slice updates as unfolded_DLASSO.py:127-140, under
    autograd -- this is what makes the reference slow (O(B*sum deg) tiny ops per iteration), so
    the timed baseline keeps it."""
    P = y.shape[1]
    out = torch.zeros_like(y)
    for b, g in enumerate(graph_list):
        for p in range(P):
            yp = y[b, p]
            for j in g.neighbors(p):
                t = yp - y[b, j]
                out[b, p] += t
                out[b, j] -= t
    return out


def reference_port_fwd_bwd(A, b, label, graph_list, param, max_param, seed: int = 7,
                           training: bool = True, backward: bool = True):
    """Loop-faithful CPU port of one training step of model #1 (forward through K iterations,
    ``loss_final.backward()``): per-agent matmul loops, Python neighbour loops, autograd.
    Returns (Y, loss_final, param.grad).  Used ONLY as bench.py's CPU baseline / reference arm."""
    P, n = A.shape[1], A.shape[3]
    AtA = atx(A, A)
    Atb = atx(A, b)
    B = b.shape[0]
    deg = degrees(graph_list, P, b.dtype)
    gen = torch.Generator().manual_seed(seed)
    y = torch.randn((B, P, n, 1), generator=gen) * 1e-2
    U = torch.randn((B, P, n, 1), generator=gen) * 1e-2
    d = torch.randn((B, P, n, 1), generator=gen) * 1e-2
    param = param.detach().clone().requires_grad_(True)
    hyp = hyp_table(param, max_param, training)
    Y = []
    for k in range(param.shape[0]):
        h = hyp[k]
        Pm = h.shape[0]
        al, ta, rh, et = (h[:, i].reshape(1, Pm, 1, 1) for i in range(4))
        a = contract(AtA, y)
        y, U, d, _ = step(a, Atb, deg, y, U, d, al, ta, rh, et, clamps_model1(k),
                          lambda v: delta_loops_autograd(graph_list, v))
        Y.append(y)
    Y = torch.stack(Y)
    _, lf = loss(Y, label)
    if not backward:              # validation pass (unfolded_train_new.py:102-128)
        return Y.detach(), lf.detach(), None
    lf.backward()
    return Y.detach(), lf.detach(), param.grad
