"""Host-side wrappers of the C ABI (``include/dadmm.h``) and the ``torch.autograd.Function``s built
on them.  PyTorch supplies device memory, streams and autograd plumbing only; every arithmetic
operation of the hot path runs in ``libdadmm_sm100.so``.

Reference lines replaced (paths relative to the reference checkout):
  contract / Contract         unfolded_DLASSO.py:69-71, :120-124  (+ autograd MmBackward)
  step_fwd / step_bwd / Step  unfolded_DLASSO.py:73-99, :127-140; gnn_dlasso_models_progressive.py:205-232
  Unfolded                    unfolded_DLASSO.py:53-109 and loss.backward() through it
  MSELoss                     gnn_dlasso_utils.py:27-88
"""
from __future__ import annotations

import ctypes as C
import math
import os
from collections import OrderedDict
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr, dtype_code, require_cuda, device_guard
from .graph import BatchGraph

INF = float("inf")


# ---------------------------------------------------------------------------------------------
# clamp schedules
# ---------------------------------------------------------------------------------------------
def clamps_model1(k: int) -> Tuple[float, float, float, float]:
    """(G, V, D, Uc) of iteration k: unfolded_DLASSO.py:80 ``max(1,30-k)``, :92 ``max(10,200-3k)``."""
    v = max(10.0, 200.0 - k * 3)
    return (max(1.0, 30.0 - k), v, INF, v)


def clamps_model3(k: int = 0) -> Tuple[float, float, float, float]:
    """gnn_dlasso_models_progressive.py:212,224,229,232: fixed 10 / 100 / 20 / 100."""
    return (10.0, 100.0, 20.0, 100.0)


def _clamps_array(clamps: Sequence[Tuple[float, float, float, float]]):
    arr = (_lib.Clamps * len(clamps))()
    for i, (g, v, d, u) in enumerate(clamps):
        arr[i] = _lib.Clamps(g, v, d, u)
    return arr


def _algo(algo) -> int:
    return _lib.ALGOS[algo] if isinstance(algo, str) else int(algo)


def same_dtype(what: str, like: torch.Tensor, **tensors) -> None:
    """Every pointer handed to the C ABI is read as ``like``'s dtype: a tensor of another dtype (an fp16 hypernetwork
    output under ``torch.autocast``, a float64 operator next to float32 states) would be read as garbage -- or, for an
    output, written out of bounds -- without any error on the device.  Raise here instead."""
    for name, t in tensors.items():
        if t is not None and t.dtype != like.dtype:
            raise _lib.DadmmError(f"{what}: `{name}` is {t.dtype} but the states are {like.dtype}; every tensor passed by "
                                  "pointer must share one dtype (float32 or float64)")


_noise_fused_ok = {}
_TORCH_RANDN = torch.randn          # the reference calls torch.randn: if a caller has replaced it, go through the replacement


def initial_noise(shape, device) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """The reference's three draws ``torch.randn(shape, device=device) * 1e-2`` (unfolded_DLASSO.py:49-51), same order.
    ``empty(shape).normal_(0, 1e-2)`` draws the same Philox numbers and applies the scale inside the generator kernel
    (x * std + 0: one fp32 rounding, as the separate multiply) -- three fewer passes over three [B,P,n] tensors per forward.
    That equivalence is an implementation property of PyTorch, so it is verified once per device and dtype on a small
    draw (bit for bit, generator state restored); where it does not hold, the two-kernel form is used."""
    dev = torch.device(device)
    if torch.randn is not _TORCH_RANDN:
        return tuple(torch.randn(shape, device=dev) * 1e-2 for _ in range(3))
    key = (str(dev), torch.get_default_dtype())
    ok = _noise_fused_ok.get(key)
    if ok is None:
        ok = False
        if dev.type == "cuda":
            with torch.random.fork_rng(devices=[dev]):
                torch.manual_seed(1234)
                a = torch.randn((4099,), device=dev) * 1e-2
                torch.manual_seed(1234)
                b = torch.empty((4099,), device=dev).normal_(0, 1e-2)
                nxt_a = None
                torch.manual_seed(1234)
                torch.randn((4099,), device=dev)
                nxt_a = torch.randn((8,), device=dev)
                torch.manual_seed(1234)
                torch.empty((4099,), device=dev).normal_(0, 1e-2)
                nxt_b = torch.randn((8,), device=dev)
                ok = bool(torch.equal(a, b) and torch.equal(nxt_a, nxt_b))
        _noise_fused_ok[key] = ok
    if ok:
        return tuple(torch.empty(shape, device=dev).normal_(0, 1e-2) for _ in range(3))
    return tuple(torch.randn(shape, device=dev) * 1e-2 for _ in range(3))


def _state3(t: torch.Tensor) -> torch.Tensor:
    """[B,P,n,1] or [B,P,n] -> contiguous [B,P,n] view/copy."""
    if t.dim() == 4:
        t = t.squeeze(-1)
    return t.contiguous()


# ---------------------------------------------------------------------------------------------
# raw ops
# ---------------------------------------------------------------------------------------------
class _ContractWsCache:
    """Workspaces of ``contract(..., constant_operator=True)`` kept between calls: the tensor-core routes leave the
    operator's scaled fp16 hi/lo copy at the front of the workspace, and while the operator (storage, version, shape), the
    batch size and the stream stay the same the next call skips the operator's two split passes
    (``dadmm_contract_prepared``).  Model #3 multiplies by the same AtA once per iteration, forward and backward
    (gnn_dlasso_models_progressive.py:158-162, :199-203): 30 of its 60 operand splits per step at configs[1]."""
    MAX = 8

    def __init__(self):
        self.entries = OrderedDict()

    def lookup(self, W, B, wsb, dev):
        key = (W.data_ptr(), W._version, tuple(W.shape), B, wsb, str(dev), torch.cuda.current_stream(dev).cuda_stream)
        ent = self.entries.get(key)
        if ent is None:
            ent = [torch.empty(wsb, dtype=torch.uint8, device=dev), False, W]     # W: its storage cannot be recycled while cached
            self.entries[key] = ent
            while len(self.entries) > self.MAX:
                self.entries.popitem(last=False)
        else:
            self.entries.move_to_end(key)
        return ent


_contract_ws = _ContractWsCache()


def contract(W: torch.Tensor, x: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: bool = False,
             algo="auto", constant_operator: bool = False) -> torch.Tensor:
    """out[b,p,:] (+)= W[p] @ x[b,p,:].  W [P,n_out,n_in], x [B,P,n_in] -> out [B,P,n_out].
    ``constant_operator``: W is a constant of the caller (in-place updates through PyTorch are noticed by its version
    counter) -- its operand copy for the tensor-core routes is kept between calls (``_ContractWsCache``)."""
    dev = require_cuda(W, x, out)
    W, x = W.contiguous(), x.contiguous()
    P, n_out, n_in = W.shape
    B = x.shape[0]
    assert x.shape == (B, P, n_in), (x.shape, W.shape)
    if out is None:
        assert not accumulate
        out = torch.empty((B, P, n_out), dtype=x.dtype, device=dev)
    assert out.is_contiguous() and out.shape == (B, P, n_out) and out.dtype == x.dtype == W.dtype
    dt, al = dtype_code(x), _algo(algo)
    with device_guard(dev):
        wsb = lib.dadmm_contract_ws_bytes(dt, al, B, P, n_out, n_in)
        if constant_operator and wsb and _op_splits.enabled:
            ent = _contract_ws.lookup(W, B, wsb, dev)
            check(lib.dadmm_contract_prepared(dt, al, B, P, n_out, n_in, ptr(W), n_out * n_in, n_in, 1, ptr(x), P * n_in, n_in, 1,
                                              ptr(out), P * n_out, n_out, 1, int(accumulate), ptr(ent[0]), wsb, int(ent[1]),
                                              stream_ptr(dev)), "dadmm_contract_prepared")
            ent[1] = True
            return out
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev) if wsb else None
        check(lib.dadmm_contract(dt, al, B, P, n_out, n_in, ptr(W), n_out * n_in, n_in, 1, ptr(x), P * n_in, n_in, 1,
                                 ptr(out), P * n_out, n_out, 1, int(accumulate), ptr(ws), wsb, stream_ptr(dev)),
              "dadmm_contract")
    return out


def observe(A: torch.Tensor, y: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """b[i,p,:] = A[p] @ y[i,:] for every agent in one launch (``set_Data``, reference gnn_data.py:13-14: a Python loop of P
    matmuls).  A [P,m,n], y [N,n] -> [N,P,m]; the shared input enters the contraction with a zero agent stride."""
    dev = require_cuda(A, y, out)
    A, y = A.contiguous(), y.contiguous()
    P, m, n = A.shape
    N = y.shape[0]
    same_dtype("observe", y, A=A, out=out)
    if out is None:
        out = torch.empty((N, P, m), dtype=y.dtype, device=dev)
    assert out.shape == (N, P, m) and out.is_contiguous() and y.shape == (N, n)
    with device_guard(dev):
        check(lib.dadmm_contract(dtype_code(y), _lib.ALGO_SIMT, N, P, m, n, ptr(A), m * n, n, 1, ptr(y), n, 0, 1,
                                 ptr(out), P * m, m, 1, 0, None, 0, stream_ptr(dev)), "dadmm_contract(observe)")
    return out


def atx(A: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """``compute_Atx`` (unfolded_DLASSO.py:120-124): Atx[:,p] = A[0,p].T @ x[:,p].
    A [1,P,m,n]; x [Bx,P,m,c] -> [Bx,P,n,c]."""
    dev = require_cuda(A, x)
    A, x = A.contiguous(), x.contiguous()
    _, P, m, n = A.shape
    Bx, c = x.shape[0], x.shape[3]
    assert x.shape[1] == P and x.shape[2] == m
    out = torch.empty((Bx, P, n, c), dtype=x.dtype, device=dev)
    dt = dtype_code(x)
    with device_guard(dev):
        for bx in range(Bx if c > 1 else 1):
            # "rows" of the contraction are the c columns of one batch entry (c>1), or the batch (c==1)
            xb, ob = (x[bx], out[bx]) if c > 1 else (x, out)
            rows = c if c > 1 else Bx
            x_sb, x_sk = (1, c) if c > 1 else (P * m, 1)
            o_sb, o_si = (1, c) if c > 1 else (P * n, 1)
            check(lib.dadmm_contract(dt, _lib.ALGO_SIMT, rows, P, n, m, ptr(A), m * n, 1, n, ptr(xb), x_sb, m * c, x_sk,
                                     ptr(ob), o_sb, n * c, o_si, 0, None, 0, stream_ptr(dev)), "dadmm_contract(Atx)")
    return out


def _hyp_struct(hyp: torch.Tensor, B: int, P: int) -> _lib.Hyp:
    """hyp [P,4] (shared, model #1 row) or [B,4,P] (per sample, model #3 layout)."""
    if hyp.dim() == 2:
        assert hyp.shape == (P, 4) and hyp.is_contiguous()
        return _lib.Hyp(hyp.data_ptr(), 0, 4, 1)
    assert hyp.shape == (B, 4, P) and hyp.is_contiguous(), hyp.shape
    return _lib.Hyp(hyp.data_ptr(), 4 * P, 1, P)


def step_fwd(graph: BatchGraph, clamps, hyp, y, U, delta, AtAy, Atb, want_delta=True, want_U=True, want_graw=True,
             flags: Optional[torch.Tensor] = None):
    """One D-ADMM iteration on [B,P,n] tensors; ``delta=None`` recomputes clampD(2L y).
    Returns (y_next, U_next|None, delta_next|None, grad_raw|None)."""
    dev = require_cuda(y, U, delta, AtAy, Atb, hyp)
    same_dtype("step_fwd", y, U=U, delta=delta, AtAy=AtAy, Atb=Atb, hyp=hyp)
    B, P, n = y.shape
    y_next = torch.empty_like(y)
    U_next = torch.empty_like(y) if want_U else None
    d_next = torch.empty_like(y) if want_delta else None
    graw = torch.empty_like(y) if want_graw else None
    cl = _lib.Clamps(*clamps)
    h = _hyp_struct(hyp, B, P)
    with device_guard(dev):
        check(lib.dadmm_step_fwd(dtype_code(y), B, P, n, C.byref(graph.c), C.byref(cl), C.byref(h), ptr(y), ptr(U),
                                 ptr(delta), ptr(AtAy), ptr(Atb), ptr(y_next), ptr(U_next), ptr(d_next), ptr(graw),
                                 ptr(flags), stream_ptr(dev)), "dadmm_step_fwd")
    return y_next, U_next, d_next, graw


def step_bwd(graph: BatchGraph, clamps, hyp, y, U, delta, graw, y_next, gy_next, gU_next, gdelta_next,
             per_sample: bool, label=None, loss_coef: float = 0.0):
    """Backward of ``step_fwd``.  Returns (gy_direct, gAtAy, gU, gdelta, ghyp) with ghyp shaped like hyp."""
    dev = require_cuda(y, U, graw, y_next)
    same_dtype("step_bwd", y, U=U, delta=delta, grad_raw=graw, y_next=y_next, gy_next=gy_next, gU_next=gU_next,
               gdelta_next=gdelta_next, hyp=hyp, label=label)
    B, P, n = y.shape
    dt = dtype_code(y)
    gy, ga, gU, gd = (torch.empty_like(y) for _ in range(4))
    part = torch.empty(lib.dadmm_partials_elems(dt, B, P, n), dtype=y.dtype, device=dev)
    cl = _lib.Clamps(*clamps)
    h = _hyp_struct(hyp, B, P)
    ghyp = torch.empty_like(hyp)
    with device_guard(dev):
        check(lib.dadmm_step_bwd(dt, B, P, n, C.byref(graph.c), C.byref(cl), C.byref(h), ptr(y), ptr(U), ptr(delta),
                                 ptr(graw), ptr(y_next), ptr(gy_next), None, ptr(gU_next), ptr(gdelta_next),
                                 ptr(label), float(loss_coef), ptr(gy), ptr(ga), ptr(gU), ptr(gd), ptr(part),
                                 stream_ptr(dev)), "dadmm_step_bwd")
        if per_sample:
            check(lib.dadmm_reduce_hyp(dt, B, P, n, ptr(part), 1, ptr(ghyp), 4 * P, 1, P, 0, stream_ptr(dev)),
                  "dadmm_reduce_hyp")
        else:
            check(lib.dadmm_reduce_hyp(dt, B, P, n, ptr(part), 0, ptr(ghyp), 0, 4, 1, 0, stream_ptr(dev)),
                  "dadmm_reduce_hyp")
    return gy, ga, gU, gd, ghyp


# ---------------------------------------------------------------------------------------------
# autograd Functions
# ---------------------------------------------------------------------------------------------
class Contract(torch.autograd.Function):
    """y [B,P,n] -> W y; backward Wt g (Wt = W^T, passed explicitly so no transpose is re-materialised)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, W, Wt, algo):
        ctx.Wt, ctx.algo = Wt, algo
        return contract(W, x, algo=algo, constant_operator=True)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g):
        return contract(ctx.Wt, g.contiguous(), algo=ctx.algo, constant_operator=True), None, None, None


class Step(torch.autograd.Function):
    """Differentiable single iteration (model #3: the hypernetwork sits between iterations).
    Inputs y, U, delta, AtAy, Atb [B,P,n]; hyp [B,4,P] or [P,4].  Outputs (y+, U+, delta+)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, y, U, delta, AtAy, Atb, hyp, graph, clamps, flags):
        y, U, delta, AtAy, Atb, hyp = (t.contiguous() for t in (y, U, delta, AtAy, Atb, hyp))
        y_next, U_next, d_next, graw = step_fwd(graph, clamps, hyp, y, U, delta, AtAy, Atb, flags=flags)
        ctx.graph, ctx.clamps = graph, clamps
        ctx.save_for_backward(y, U, delta, graw, y_next, hyp)
        return y_next, U_next, d_next

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gy_next, gU_next, gd_next):
        y, U, delta, graw, y_next, hyp = ctx.saved_tensors
        c = lambda t: None if t is None else t.contiguous().to(y.dtype)
        gy, ga, gU, gd, ghyp = step_bwd(ctx.graph, ctx.clamps, hyp, y, U, delta, graw, y_next, c(gy_next), c(gU_next),
                                        c(gd_next), per_sample=(hyp.dim() == 3))
        return gy, gU, gd, ga, None, ghyp, None, None, None


def solver_operators(cache: dict, A: torch.Tensor, device, dtype):
    """(A, AtA [P,n,n], AtA^T, A^T [P,n,m]) on ``device`` in the dtype of the states, for the modules' ``A`` attribute
    (reference unfolded_DLASSO.py:12-16: ``A`` is a plain attribute, ``AtA`` is computed from it once).  One entry per
    device, validated against the identity and version of ``A`` -- assigning a new ``model.A`` or writing to it in place
    rebuilds the operators (and frees the previous ones); the states are drawn by ``torch.randn`` in the default dtype, so
    an ``A`` of another dtype is cast to it here, as every tensor handed to the library must share one dtype."""
    dev = torch.device(device)
    key = (dtype, id(A), A._version)
    hit = cache.get(str(dev))
    if hit is None or hit[0] != key:
        Ad = A.detach().to(device=dev, dtype=dtype)
        W = atx(Ad, Ad)[0].contiguous()                 # AtA_p = A_p^T A_p   (reference :16)
        Wt = W.transpose(1, 2).contiguous()
        if torch.equal(W, Wt):
            Wt = W
        At = Ad[0].transpose(1, 2).contiguous()         # [P,n,m]: K-major operator for Atb on the tensor-core path
        hit = (key, (Ad, W, Wt, At))
        cache[str(dev)] = hit
    return hit[1]


def clear_caches() -> None:
    """Release the process-wide caches: persistent operator splits (``_OpSplitCache``) and ingested graph batches
    (``graph._cache``).  The modules' own operator caches (``model._ops``) go with the module."""
    from . import graph as _graph
    _op_splits.entries.clear()
    _contract_ws.entries.clear()
    _graph._cache.clear()


def _factor_struct(factor, P, n, like):
    """(Factor struct, tensors kept alive) for a factor (F1, F2[, rhs]), or None.  Shapes are validated here: a wrong
    factor would be a silent wrong answer on the device."""
    if factor is None:
        return None
    F1, F2 = (t.contiguous() for t in factor[:2])
    rhs = factor[2].contiguous() if len(factor) > 2 and factor[2] is not None else None
    require_cuda(F1, F2, rhs)
    m = F1.shape[-2]
    if tuple(F1.shape[-3:]) != (P, m, n) or tuple(F2.shape[-3:]) != (P, n, m) or F1.dtype != like.dtype or F2.dtype != like.dtype:
        raise _lib.DadmmError(f"factor shapes {tuple(F1.shape)}, {tuple(F2.shape)} do not describe a [P={P},n={n},n] operator")
    if rhs is not None and (tuple(rhs.shape) != (like.shape[0], P, m) or rhs.dtype != like.dtype):
        raise _lib.DadmmError(f"factor rhs shape {tuple(rhs.shape)} is not [B={like.shape[0]},P={P},m={m}]")
    return _lib.Factor(int(m), ptr(F1), ptr(F2), ptr(rhs)), (F1, F2, rhs)


class _OpSplitCache:
    """Persistent operator splits (``dadmm_op_split``).  The reference's operator is a constructor-time constant
    (unfolded_DLASSO.py:12-16), so the scaled fp16 hi/lo copies the tensor-core contraction reads are made once per
    (operator storage, version, route, stream) instead of by every forward pass and every reverse sweep.  An in-place
    update of the operator bumps its version counter and invalidates the entry; ``DADMM_OP_SPLIT_CACHE=0`` disables it."""
    MAX = 4

    def __init__(self):
        self.entries = OrderedDict()
        self.enabled = os.environ.get("DADMM_OP_SPLIT_CACHE", "1") != "0"

    def lookup(self, dev, dt, al, B, P, n, W, fac):
        """-> [buffer, ready, keepalive] or None when the shape does not take the fused tensor-core path."""
        if not self.enabled:
            return None
        m = int(fac[0].m) if fac else 0
        nbytes = int(lib.dadmm_unfolded_op_split_bytes(dt, al, B, P, n, m))
        if nbytes == 0:
            return None
        two = bool(m and lib.dadmm_unfolded_uses_factor(dt, al, B, P, n, m))
        src = tuple(fac[1][:2]) if two else (W,)
        key = (two, nbytes, str(dev), torch.cuda.current_stream(dev).cuda_stream,
               tuple((t.data_ptr(), t._version, tuple(t.shape), tuple(t.stride())) for t in src))
        ent = self.entries.get(key)
        if ent is None:
            ent = [torch.empty(nbytes, dtype=torch.uint8, device=dev), False, src]      # src: the storage cannot be recycled while cached
            self.entries[key] = ent
            while len(self.entries) > self.MAX:
                self.entries.popitem(last=False)
        else:
            self.entries.move_to_end(key)
        return ent

    @staticmethod
    def struct(ent):
        return None if ent is None else _lib.OpSplit(ent[0].data_ptr(), ent[0].numel(), int(ent[1]))


_op_splits = _OpSplitCache()


class Unfolded(torch.autograd.Function):
    """K fused iterations of model #1: hyp [K,P,4] -> Y [K,B,P,n].  Saves Y, U_k and the raw
    gradients r_k; backward runs the reverse sweep on device and returns d/d hyp only (the
    reference has no other differentiable input on this path)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, hyp, W, Wt, Atb, y0, U0, d0, graph, clamps, algo, flags, handle, factor=None, factor_t=None):
        """factor = (F1 [P,m,n], F2 [P,n,m][, rhs [B,P,m]]) with W = F2 F1 (and Atb = F2 rhs), factor_t likewise for Wt
        (``dadmm_factor``): lets the library evaluate the contraction in two stages when that is cheaper
        (AtA = A^T A: F1 = A, F2 = A^T, rhs = b)."""
        dev = require_cuda(hyp, W, Wt, Atb, y0, U0, d0)
        same_dtype("Unfolded", y0, hyp=hyp, W=W, Wt=Wt, Atb=Atb, U0=U0, d0=d0)
        if handle is not None:
            handle.pending = None         # an offer no reverse sweep collected belongs to an earlier forward pass
        hyp, y0, U0, d0 = (t.contiguous() for t in (hyp, y0, U0, d0))
        Atb = Atb.contiguous() if Atb is not None else None      # None: only with a factor rhs on the two-stage route
        K, P, _ = hyp.shape
        B, _, n = y0.shape
        dt, al = dtype_code(y0), _algo(algo)
        need_grad = ctx.needs_input_grad[0]
        Y = torch.empty((K, B, P, n, 1), dtype=y0.dtype, device=dev)
        U_save = torch.empty((max(K - 1, 1), B, P, n), dtype=y0.dtype, device=dev) if need_grad else None
        R_save = torch.empty((K, B, P, n), dtype=y0.dtype, device=dev) if need_grad else None
        carr = _clamps_array(clamps)
        fac, fac_t = _factor_struct(factor, P, n, y0), _factor_struct(factor_t, P, n, y0)
        sums = None
        if handle is not None and K >= 3 and y0.dtype == torch.float32:
            # label-free loss sums (dadmm_loss_sums): lets compute_loss skip its full read of Y for the iterations the
            # library reports as valid
            S = torch.empty((K, B, n), dtype=y0.dtype, device=dev)
            sq = torch.empty(K, dtype=torch.float64, device=dev)
            valid = (C.c_int32 * K)()
            sums = (_lib.LossSums(ptr(S), ptr(sq), valid), S, sq, valid)
        with device_guard(dev):
            wsb = lib.dadmm_unfolded_ws_bytes(dt, al, B, P, n, K, 0, fac[0].m if fac else 0)
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            W = W.contiguous()
            ent = _op_splits.lookup(dev, dt, al, B, P, n, W, fac)
            ops = _OpSplitCache.struct(ent)
            check(lib.dadmm_unfolded_fwd(dt, al, B, P, n, K, C.byref(graph.c), carr, ptr(hyp), ptr(W),
                                         C.byref(fac[0]) if fac else None, ptr(Atb), ptr(y0),
                                         ptr(U0), ptr(d0), ptr(Y), ptr(U_save), ptr(R_save), ptr(ws), wsb, ptr(flags),
                                         C.byref(sums[0]) if sums else None, C.byref(ops) if ops else None,
                                         stream_ptr(dev)), "dadmm_unfolded_fwd")
            if ent is not None:
                ent[1] = True
        if sums is not None and any(sums[3]):
            handle.sums = (sums[1], sums[2], [bool(v) for v in sums[3]], Y.data_ptr())
        if need_grad:
            ctx.save_for_backward(hyp, Wt, y0, U0, d0, Y, U_save, R_save)
            ctx.graph, ctx.clamps, ctx.algo = graph, carr, al
            ctx.handle = handle
            ctx.fac_t = fac_t            # (struct, tensors kept alive)
        return Y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gY):
        hyp, Wt, y0, U0, d0, Y, U_save, R_save = ctx.saved_tensors
        dev = Y.device
        K, B, P, n = Y.shape[:4]
        dt = dtype_code(Y)
        label, coef, coef_dev = None, None, None
        fused = ctx.handle.take() if ctx.handle is not None else None
        if fused is not None:
            label, coefs, sentinel = fused
            if isinstance(coefs, torch.Tensor):
                coef_dev = coefs.to(device=dev, dtype=torch.float64).contiguous()      # stays on the device: no host sync
            else:
                coef = (C.c_double * K)(*coefs)
            if gY is not None and gY.data_ptr() == sentinel.data_ptr():
                gY = None       # zero placeholder emitted by MSELoss.backward: gradient arrives through (label, coef)
        if gY is not None:
            gY = gY.contiguous().to(Y.dtype)
        ghyp = torch.empty_like(hyp)
        fac_t = ctx.fac_t
        with device_guard(dev):
            wsb = lib.dadmm_unfolded_ws_bytes(dt, ctx.algo, B, P, n, K, 1, fac_t[0].m if fac_t else 0)
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            Wt = Wt.contiguous()
            ent = _op_splits.lookup(dev, dt, ctx.algo, B, P, n, Wt, fac_t) if K > 1 else None
            ops = _OpSplitCache.struct(ent)
            check(lib.dadmm_unfolded_bwd(dt, ctx.algo, B, P, n, K, C.byref(ctx.graph.c), ctx.clamps, ptr(hyp), ptr(Wt),
                                         C.byref(fac_t[0]) if fac_t else None, ptr(y0), ptr(U0), ptr(d0), ptr(Y), ptr(U_save), ptr(R_save), ptr(gY),
                                         ptr(label), coef, ptr(coef_dev), ptr(ghyp), ptr(ws), wsb, C.byref(ops) if ops else None,
                                         stream_ptr(dev)), "dadmm_unfolded_bwd")
            if ent is not None:
                ent[1] = True
        return (ghyp,) + (None,) * 13


def tc_linear_ok(M: int, c_in: int, c_out: int) -> bool:
    """Do both products of a bias-free linear layer on [M, c_in] rows -- x W^T and g W -- take the tensor-core contraction?"""
    tc = lib.dadmm_contract_uses_tensor_cores
    return bool(tc(_lib.F32, _lib.ALGO_AUTO, M, 1, c_out, c_in)) and bool(tc(_lib.F32, _lib.ALGO_AUTO, M, 1, c_in, c_out))


class TCLinear(torch.autograd.Function):
    """y = x W^T for x [M, c_in], W [c_out, c_in] (``nn.Linear`` layout) on the library's fp32-accurate tensor-core
    contraction (tcgen05, scaled fp16 hi/lo operand pairs: the accuracy of an fp32 FMA loop at several times cuBLAS's fp32
    SGEMM rate): the dense products of the model-#3 hypernetwork (gnn_dlasso_models_progressive.py:52-68, :93-106), which
    PyTorch evaluates on the FP32 FMA pipe.  Backward: g W through the same kernel; g^T x (a reduction over the M rows) by
    cuBLAS."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, W):
        ctx.save_for_backward(x, W)
        return contract(W.unsqueeze(0), x.unsqueeze(1)).squeeze(1)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g):
        x, W = ctx.saved_tensors
        g = g.contiguous()
        gx = contract(W.t().contiguous().unsqueeze(0), g.unsqueeze(1)).squeeze(1) if ctx.needs_input_grad[0] else None
        gW = g.t().mm(x) if ctx.needs_input_grad[1] else None
        return gx, gW


class SplitKLinear(torch.autograd.Function):
    """y = x W^T for tall x [M, c_in] (M = B*P rows of the hypernetwork's GCN layers), products by cuBLAS.  The weight
    gradient g^T x reduces over the M rows; as ONE GEMM its output has (c_out/64)(c_in/64) ~ 49 tiles for 148 SMs and a
    5120-long k loop (52 us per call at configs[1], a third of the step's GEMM time: torch-profiler view in
    profiles/r02_model3_torch_profiler.txt).  Here the reduction is cut into SPLIT batched products summed afterwards."""
    SPLIT = 8

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, W):
        ctx.save_for_backward(x, W)
        return x.matmul(W.t())

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g):
        x, W = ctx.saved_tensors
        lead = x.shape[:-1]
        x2, g2 = x.reshape(-1, x.shape[-1]), g.reshape(-1, g.shape[-1])
        gx = g2.matmul(W).reshape(*lead, W.shape[1]) if ctx.needs_input_grad[0] else None
        gW = None
        if ctx.needs_input_grad[1]:
            M, S = x2.shape[0], SplitKLinear.SPLIT
            if M >= 2048 and M % S == 0:
                gW = torch.bmm(g2.reshape(S, M // S, -1).transpose(1, 2), x2.reshape(S, M // S, -1)).sum(dim=0)
            else:
                gW = g2.t().matmul(x2)
        return gx, gW


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``F.linear`` with the product on the tensor-core contraction wherever the shape takes it (CUDA, float32, no
    autocast); otherwise PyTorch's own."""
    lead, c_in = x.shape[:-1], x.shape[-1]
    M = int(math.prod(lead)) if lead else 1
    if (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and not torch.is_autocast_enabled()
            and tc_linear_ok(M, c_in, weight.shape[0])):
        y = TCLinear.apply(x.reshape(M, c_in).contiguous(), weight).reshape(*lead, weight.shape[0])
        return y if bias is None else y + bias
    return torch.nn.functional.linear(x, weight, bias)


class GCNEpilogue(torch.autograd.Function):
    """Everything a graph-convolution layer of the model-#3 hypernetwork does after its dense product, for the whole
    batch in one kernel each way (``dadmm_gcn_epilogue_fwd`` / ``_bwd``; reference gnn_dlasso_models_progressive.py:37-72,
    there a Python loop over the samples):

        Z = A_hat_b H_b + bias;  A = leaky_relu(Z);  BatchNorm over the P nodes of problem b;  out = BN(A) * mask

    H [B,P,C] (= ``conv.lin(x)``), adj [B,P,P], mask [B,P,C] or None (dropout keep mask already divided by 1-p).  Returns
    (out, mean, var): the per-problem batch statistics ([B,C], biased variance; training mode only, else None) feed the
    running-statistics update and carry no gradient.  Gradients: H, conv bias, BN weight and bias."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, H, adj, bias, bn_w, bn_b, run_mean, run_var, training, eps, slope, mask):
        dev = require_cuda(H, adj, bias, bn_w, bn_b, mask)
        if H.dtype != torch.float32:
            raise _lib.DadmmError("GCNEpilogue computes in float32")
        same_dtype("GCNEpilogue", H, adj=adj, bias=bias, bn_w=bn_w, bn_b=bn_b, mask=mask, run_mean=run_mean, run_var=run_var)
        B, P, Cc = H.shape
        H, adj = H.contiguous(), adj.contiguous()
        mask = mask.contiguous() if mask is not None else None
        out, act = torch.empty_like(H), torch.empty_like(H)
        mean = torch.empty((B, Cc), dtype=H.dtype, device=dev) if training else None
        var = torch.empty((B, Cc), dtype=H.dtype, device=dev) if training else None
        with device_guard(dev):
            check(lib.dadmm_gcn_epilogue_fwd(B, P, Cc, ptr(H), ptr(adj), ptr(bias), ptr(bn_w), ptr(bn_b), ptr(run_mean), ptr(run_var),
                                             int(bool(training)), float(eps), float(slope), ptr(mask), ptr(out), ptr(act), ptr(mean),
                                             ptr(var), stream_ptr(dev)), "dadmm_gcn_epilogue_fwd")
        # training mode never reads the running statistics (and the caller updates them in place right after this call)
        ctx.save_for_backward(adj, bn_w, None if training else run_mean, None if training else run_var, mask, act, mean, var)
        ctx.cfg = (bool(training), float(eps), float(slope))
        if training:
            ctx.mark_non_differentiable(mean, var)
        ctx.set_materialize_grads(False)
        return out, mean, var

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gout, _gm, _gv):
        adj, bn_w, run_mean, run_var, mask, act, mean, var = ctx.saved_tensors
        if gout is None:
            return (None,) * 11
        training, eps, slope = ctx.cfg
        B, P, Cc = act.shape
        dev = act.device
        gout = gout.contiguous().to(act.dtype)
        gH = torch.empty_like(act)
        rows = int(lib.dadmm_gcn_partial_rows(B, Cc))
        part = torch.empty((rows, 3, Cc), dtype=act.dtype, device=dev)
        with device_guard(dev):
            check(lib.dadmm_gcn_epilogue_bwd(B, P, Cc, ptr(gout), ptr(adj), ptr(bn_w), ptr(run_mean), ptr(run_var), int(training),
                                             eps, slope, ptr(mask), ptr(act), ptr(mean), ptr(var), ptr(gH), ptr(part), stream_ptr(dev)),
                  "dadmm_gcn_epilogue_bwd")
        sums = part.sum(dim=0)                  # rows of per-CTA partial sums -> (d bn_w, d bn_b, d bias)
        return gH, None, sums[2], sums[0], sums[1], None, None, None, None, None, None


def loss_per_iteration(Y: torch.Tensor, label: torch.Tensor, B_norm: Optional[int] = None, handle=None) -> torch.Tensor:
    """losses[k] = sum (Y[k]-label)^2 / (P*B_norm*n)  (gnn_dlasso_utils.py:54-66).  No autograd.

    When ``handle`` carries the label-free sums the forward pass left behind (``dadmm_loss_sums``), the inner iterations
    are evaluated from them -- a read of [K,B,n] instead of [K,B,P,n] -- and only the first and the last iteration (the
    one the drivers back-propagate) read Y."""
    dev = require_cuda(Y, label)
    K, B, P, n = Y.shape[:4]
    Yc, lab = Y.contiguous(), label.contiguous()
    dt = dtype_code(Yc)
    Bn = int(B_norm or B)
    losses = torch.empty(K, dtype=Y.dtype, device=dev)
    sums = getattr(handle, "sums", None)
    if sums is not None and (sums[3] != Yc.data_ptr() or len(sums[2]) != K):
        sums = None
    with device_guard(dev):
        wsb = lib.dadmm_loss_ws_bytes(dt, K, B, P, n)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)

        def exact(k0, k1):
            es = Yc.element_size()
            check(lib.dadmm_loss_fwd(dt, k1 - k0, B, P, n, Bn, C.c_void_p(Yc.data_ptr() + k0 * B * P * n * es), ptr(lab),
                                     C.c_void_p(losses.data_ptr() + k0 * es), ptr(ws), wsb, stream_ptr(dev)), "dadmm_loss_fwd")
        if sums is None:
            exact(0, K)
        else:
            S, sq, valid, _ = sums
            use = list(valid)
            use[K - 1] = False                 # the final iteration is always evaluated exactly from Y
            k = 0
            while k < K:                       # maximal runs of equal kind
                j = k
                while j < K and use[j] == use[k]:
                    j += 1
                if use[k]:
                    check(lib.dadmm_loss_from_sums(dt, k, j, B, P, n, Bn, ptr(S), ptr(sq), ptr(lab), ptr(Yc), ptr(losses), ptr(ws), wsb,
                                                   stream_ptr(dev)), "dadmm_loss_from_sums")
                else:
                    exact(k, j)
                k = j
    return losses


class FusedLossHandle:
    """Side channel between ``MSELoss.backward`` and ``Unfolded.backward`` (same autograd pass)."""

    def __init__(self):
        self.pending = None
        self.sums = None          # (agent_sum [K,B,n], sumsq [K], valid [K], Y.data_ptr()) left by Unfolded.forward

    def offer(self, label, coefs, sentinel) -> bool:
        if self.pending is not None:
            return False
        self.pending = (label, coefs, sentinel)
        return True

    def take(self):
        p, self.pending = self.pending, None
        return p


class MSELoss(torch.autograd.Function):
    """Per-iteration losses [K] with a fused backward.

    When Y comes straight from ``Unfolded`` (``Y._dadmm_handle`` set by the module), backward hands the
    gradient to the reverse sweep as (label, coef[k]) -- the kernel synthesises
    gY[k] = coef[k]*(Y[k]-label) on the fly -- and returns a stride-0 zero placeholder instead of a
    dense [K,B,P,n] tensor.  Otherwise it materialises gY with ``dadmm_loss_bwd``.

    Restriction of the fused route: the gradient with respect to Y itself is not materialised, so anything that observes
    it -- ``Y.retain_grad()``, a hook on Y, ``torch.autograd.grad(loss, Y)`` -- would see zeros.  ``Y.retain_grad()`` and
    hooks registered before ``compute_loss`` are detected and take the dense route; for ``autograd.grad(loss, Y)`` delete
    ``Y._dadmm_handle`` first.  An offer the reverse sweep never collected (a backward pass that stopped at Y) is dropped
    by the next forward pass or loss evaluation, never applied to a later one."""

    @staticmethod
    def forward(ctx, Y, label, B_norm, handle):
        same_dtype("MSELoss", Y, label=label)
        ctx.save_for_backward(Y, label)
        if handle is not None and (Y.retains_grad or Y._backward_hooks):
            handle = None                 # someone wants to see d loss / d Y: dense route
        ctx.B_norm, ctx.handle = B_norm, handle
        if handle is not None:
            handle.pending = None         # a (label, coef) offer is valid for the backward pass of THIS loss only
        return loss_per_iteration(Y, label, B_norm, handle)

    @staticmethod
    def backward(ctx, g_losses):
        Y, label = ctx.saved_tensors
        K, B, P, n = Y.shape[:4]
        scale = 2.0 / (P * (ctx.B_norm or B) * n)
        if ctx.handle is not None:
            # coefficients stay on the device (dadmm_unfolded_bwd: loss_coef_dev): reading them here would stall the host
            # until the whole forward pass has drained, once per training step
            coefs_dev = g_losses.detach().to(torch.float64) * scale
            sentinel = torch.zeros((), dtype=Y.dtype, device=Y.device).expand(Y.shape)
            if ctx.handle.offer(label.contiguous(), coefs_dev, sentinel):
                return sentinel, None, None, None
        if torch.cuda.is_current_stream_capturing():
            # under CUDA-graph capture the coefficients cannot visit the host: dense gY = coef[k] (Y[k] - label) in one pass
            c = (g_losses.detach().to(Y.dtype) * scale).view(K, 1, 1, 1, 1)
            return (Y - label.view(1, B, 1, Y.shape[3], 1)) * c, None, None, None
        coefs = [float(v) * scale for v in g_losses.detach().to("cpu", torch.float64).tolist()]
        gY = torch.empty_like(Y, memory_format=torch.contiguous_format)
        with device_guard(Y.device):
            check(lib.dadmm_loss_bwd(dtype_code(Y), K, B, P, n, ptr(Y.contiguous()), ptr(label.contiguous()),
                                     (C.c_double * K)(*coefs), ptr(gY), stream_ptr(Y.device)), "dadmm_loss_bwd")
        return gY, None, None, None
