"""Drop-in replacement of the reference module ``unfolded_DLASSO.py`` (same class names, constructor /
``forward`` signatures, attributes and state-dict keys) whose hot loop runs in ``libdadmm_sm100.so``.

Reference behaviour mirrored (paths relative to the reference checkout):
  * ``DLASSO_unfolded.__init__``  unfolded_DLASSO.py:10-32
  * ``DLASSO_unfolded.forward``   unfolded_DLASSO.py:34-110  -> one fused K-iteration autograd.Function
  * ``seq_hyperparam``            unfolded_DLASSO.py:148-168 (stays PyTorch: it is the learnable
    parameter; the kernels take its [K,P,4] table and return d loss / d table)
The per-sample Python loops of ``compute_sum_neighbors`` / ``compute_delta`` are replaced by a one-off
CSR build (``dadmm_b200.graph``) and the in-kernel consensus operator.  There is no CPU path:
``forward`` raises when ``b`` is not a CUDA tensor.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from dadmm_b200 import functional as DF
from dadmm_b200.graph import BatchGraph


class _PrefixSums(torch.autograd.Function):
    """s[k] = torch.sum(param[:k+1], dim=0), k < K  (unfolded_DLASSO.py:157).  Forward keeps the reference's reduction call
    by call, so the rows agree bit for bit; backward is the closed form d param[i] = sum_{k >= i} d s[k] as one masked
    product and one reduction instead of the ~3K tiny slice-backward / accumulate kernels autograd chains up behind K
    slices -- at small per-GPU batches those launches were a measurable share of the step."""

    _masks = {}

    @staticmethod
    def forward(ctx, param, K):
        ctx.rows, ctx.K = param.shape[0], K
        return torch.stack([torch.sum(param[:k + 1], dim=0) for k in range(K)])

    @staticmethod
    def backward(ctx, gs):
        K = ctx.K
        key = (K, gs.device, gs.dtype)
        mask = _PrefixSums._masks.get(key)
        if mask is None:
            mask = torch.triu(torch.ones((K, K), dtype=gs.dtype, device=gs.device))      # mask[i, k] = 1 for k >= i
            _PrefixSums._masks[key] = mask
        g = (mask.reshape(K, K, *([1] * (gs.dim() - 1))) * gs.unsqueeze(0)).sum(dim=1)
        if ctx.rows > K:
            g = torch.cat([g, g.new_zeros((ctx.rows - K,) + tuple(g.shape[1:]))])
        return g, None


class seq_hyperparam(nn.Module):
    """Learned per-iteration hyper-parameters (alpha, tau, rho, eta): ``param`` [K, P|1, 4], zero-init."""

    def __init__(self, hyp_shape, max_param, args=None):
        super().__init__()
        self.param = nn.Parameter(torch.zeros(hyp_shape))
        self.max_param = max_param.unsqueeze(0)
        self.args = args

    def _max_param_on(self, device):
        """``max_param`` is a plain attribute in the reference (it stays on the CPU when the module moves): keep one copy
        per device instead of a host-to-device copy in every forward; replaced or edited in place, it is copied again."""
        src = self.max_param
        if src.device == device:
            return src
        key, hit = (id(src), src._version, str(device)), self.__dict__.get("_max_param_dev")
        if hit is None or hit[0] != key:
            hit = (key, src.to(device))
            self.__dict__["_max_param_dev"] = hit
        return hit[1]

    def _squash(self, s):
        # s: [..., P|1, 4] pre-activation -> sigmoid * max, train-mode penalty, clamp  (:158-167)
        h = torch.sigmoid(s) * self._max_param_on(s.device)
        if self.training and self.args is not None:
            mean = h.sum(dim=(-2, -1), keepdim=True) / (h.shape[-2] * h.shape[-1])
            h = torch.where(mean > self.args.max_penalty_threshold, h * self.args.penalty_reduction_factor, h)
        return torch.clamp(h, min=1e-4, max=0.99)

    def forward(self, k):
        """Hyper-parameters of iteration k, [P|1, 4, 1] (same values as the reference's forward(k))."""
        s = torch.sum(self.param[:k + 1], dim=0)
        return self._squash(s.reshape(-1, 4)).unsqueeze(-1)

    def table(self, K):
        """All K rows at once, [K, P|1, 4]: row k == forward(k).squeeze(-1).  The running sum uses the same
        ``torch.sum(param[:k+1])`` reduction as the reference so the rows agree bit for bit."""
        return self._squash(_PrefixSums.apply(self.param, K))


class DLASSO_unfolded(nn.Module):
    def __init__(self, A, args):
        super().__init__()
        self.A = A                                      # [1, P, m, n], plain attribute as in the reference
        _, self.P, self.m, self.n = self.A.shape
        self.K = args.GHN_iter_num
        self.DADMM_mode = args.DADMM_mode
        hyp_shape = [self.K, 1 if args.DADMM_mode == 'same' else self.P, 4]
        max_param = torch.tensor([args.alpha_max, args.tau_max, args.rho_max, args.eta_max], device=A.device)
        self.seq_hyp = seq_hyperparam(hyp_shape, max_param, args)
        self.max_param = max_param.unsqueeze(0)
        self.args = args
        # knobs of the B200 path
        self.contract_algo = "auto"      # "auto" | "simt" | "tc" (tcgen05 3xTF32)
        # True: one device flag read per forward (reference: 4 host syncs / iteration) and the reference's reset / skip
        # semantics on a hit; "deferred": the flags accumulate in a sticky device buffer that only nonfinite_seen() reads
        # (no host sync in forward: what CUDA-graph capture of a whole step needs, dadmm_b200/graphs.py); False: no flags
        self.check_finite = True
        self.two_stage = True            # offer AtA = A^T A to the library as a factor pair (used where 4mn << 2n^2)
        self._ops = {}                   # device -> (A, AtA [P,n,n], AtA^T [P,n,n])

    # ------------------------------------------------------------------ operators
    def _operators(self, device, dtype=None):
        return DF.solver_operators(self._ops, self.A, device, dtype or torch.get_default_dtype())

    @property
    def AtA(self):
        """[1,P,n,n] on ``A``'s device (computed by the CUDA contraction kernel on first use)."""
        dev = self.A.device if self.A.is_cuda else torch.device("cuda", torch.cuda.current_device())
        return self._operators(dev)[1].unsqueeze(0).to(self.A.device)

    # ------------------------------------------------------------------ forward
    def forward(self, b, graph_list, K=None, noise=None):
        """b [B,P,m,1]; graph_list: B graphs over nodes 0..P-1.  Returns (Y [K,B,P,n,1], hyp [P|1,4,1]).

        ``noise`` (not in the reference): the initial (y, U, delta), each [B,P,n,1], in place of the three ``randn * 1e-2``
        draws -- a rank of a batch-sharded job passes its slice of the full batch's draws (``dadmm_b200.dist.sharded_noise``)
        so that N ranks reproduce the single-process run."""
        if len(b) != len(graph_list):
            raise ValueError(f"len(b)={len(b)} != len(graph_list)={len(graph_list)}")
        batch_size, device = len(b), b.device
        K = self.K if K is None else min(K, self.K)
        DF.require_cuda(b)
        A, W, Wt, At = self._operators(device)
        # A_p^T b_p (reference :45) -- not needed when the library forms the residual as A^T (A y - b) from the factor pair
        Atb = None if self._residual_from_factor(W, batch_size) else self._atb(At, b, W.dtype)
        graph = BatchGraph.from_graph_list(graph_list, self.P, device)
        # initial noise: same three draws, same order / shape / device as the reference (:49-51)
        if noise is None:
            # torch.randn(...) * 1e-2 three times, in the reference's order (:49-51); scale applied by the generator kernel
            y0, U0, d0 = DF.initial_noise((batch_size, self.P, self.n, 1), device)
        else:
            y0, U0, d0 = (t.to(device=device, dtype=W.dtype).reshape(batch_size, self.P, self.n, 1) for t in noise)
        table = self.seq_hyp.table(K)                                   # [K, P|1, 4]
        Y = self._run(table, W, Wt, Atb, y0, U0, d0, graph, K, b)
        return Y, table[K - 1].unsqueeze(-1)

    def _atb(self, At, b, dtype):
        return DF.contract(At, b.to(dtype).squeeze(-1), algo=self.contract_algo)

    def _residual_from_factor(self, W, batch_size):
        """True when ``_run`` will hand (A, A^T, b) to the library AND the library takes the two-stage route for this
        shape (``dadmm_unfolded_uses_factor``): the forward then never reads ``Atb``."""
        if not (self.two_stage and W.dtype == torch.float32 and getattr(self, "two_stage_rhs", True)):
            return False
        return bool(DF.lib.dadmm_unfolded_uses_factor(DF.dtype_code(W), DF._algo(self.contract_algo), batch_size, self.P,
                                                      self.n, self.m))

    def _run(self, table, W, Wt, Atb, y0, U0, d0, graph, K, b=None):
        hyp = table.expand(K, self.P, 4).contiguous().to(W.dtype)
        clamps = [DF.clamps_model1(k) for k in range(K)]
        deferred = self.check_finite == "deferred"
        if deferred:
            flags = self._sticky_flags(W.device, K)
        else:
            flags = torch.zeros(K, dtype=torch.int32, device=W.device) if self.check_finite else None
        handle = DF.FusedLossHandle()
        factor = factor_t = None
        if self.two_stage and W.dtype == torch.float32:
            A, _, _, At = self._operators(W.device)
            factor_t = (A[0], At)                        # AtA y = A^T (A y); AtA is symmetric, so the same pair serves backward
            rhs = b.to(W.dtype).squeeze(-1) if (b is not None and getattr(self, "two_stage_rhs", True)) else None
            factor = (A[0], At, rhs)                     # forward: residual A^T (A y - b) = AtA y - Atb
        Y = DF.Unfolded.apply(hyp, W, Wt, Atb, y0.squeeze(-1), U0.squeeze(-1), d0.squeeze(-1), graph, clamps,
                              self.contract_algo, flags, handle, factor, factor_t)
        if flags is not None and not deferred and bool(flags.any()):
            # non-finite values seen: redo the batch on the guarded path, which reproduces the reference's
            # reset / skip semantics (:55-61, :84-86, :102-104) iteration by iteration
            if Atb is None:
                Atb = self._atb(self._operators(W.device)[3], b, W.dtype)
            return self._run_guarded(hyp, W, Wt, Atb, y0, U0, d0, graph, K)
        Y._dadmm_handle = handle
        if flags is not None and not deferred:
            Y._dadmm_finite = Y._version      # no kernel saw a non-finite value => Y is finite (compute_loss skips its scan)
        return Y

    def _sticky_flags(self, device, K):
        buf = self.__dict__.get("_flag_buf")
        if buf is None or buf.device != device or buf.numel() < K:
            buf = torch.zeros(max(K, self.K), dtype=torch.int32, device=device)
            self.__dict__["_flag_buf"] = buf
        return buf

    def nonfinite_seen(self, reset=True):
        """``check_finite = "deferred"``: has any forward pass since the last call met a non-finite value?  (One host read;
        the kernels only ever OR bits into the buffer.)  Such a pass ran the plain recurrence on non-finite data -- rerun
        the batch with ``check_finite = True`` for the reference's reset / skip semantics (:55-61, :84-86, :102-104)."""
        buf = self.__dict__.get("_flag_buf")
        if buf is None:
            return False
        hit = bool(buf.any())
        if reset and hit:
            buf.zero_()
        return hit

    def _run_guarded(self, hyp, W, Wt, Atb, y0, U0, d0, graph, K):
        bad = lambda t: bool(torch.isnan(t).any() or torch.isinf(t).any())
        y, U, d = y0.squeeze(-1), U0.squeeze(-1), d0.squeeze(-1)
        Y = []
        for k in range(K):
            if bad(y):
                print(f"Warning: NaN/Inf detected in y_k at iteration {k}, resetting...")
                y = torch.zeros_like(y)
            if bad(U):
                print(f"Warning: NaN/Inf detected in U_k at iteration {k}, resetting...")
                U = torch.zeros_like(U)
            a = DF.Contract.apply(y, W, Wt, self.contract_algo)
            clamps = DF.clamps_model1(k)
            flag = torch.zeros(1, dtype=torch.int32, device=y.device)
            y_n, U_n, d_n = DF.Step.apply(y, U, d, a, Atb, hyp[k], graph, clamps, flag)
            if int(flag) & 4:
                print(f"Warning: NaN/Inf in gradient at iteration {k}, skipping update...")
                # grad := 0  =>  y_next = clamp(y_k - alpha_k * 0); consensus / dual update as usual (:86-99).  The product is
                # kept: a non-finite alpha_k makes y_next non-finite here exactly as in the reference (-> the :102-104 guard)
                y_n = torch.clamp(y - hyp[k][:, 0].reshape(1, -1, 1) * torch.zeros_like(y), -clamps[1], clamps[1])
                d_n = self.compute_delta(None, y_n.unsqueeze(-1), _graph=graph).squeeze(-1)
                U_n = torch.clamp(U + d_n * hyp[k][:, 3].reshape(1, -1, 1), -clamps[3], clamps[3])
            if bad(y_n):
                print(f"Warning: NaN/Inf in y_next at iteration {k}, using previous value...")
                y_n = y
            y, U, d = y_n, U_n, d_n
            Y.append(y)
        return torch.stack(Y).unsqueeze(-1)

    # ------------------------------------------------------------------ reference helper API
    def compute_sum_neighbors(self, graph_list, device):
        """[B,P,1,1] float degrees (reference :111-118)."""
        host = BatchGraph.build_host(graph_list, self.P)
        deg = torch.from_numpy(host.deg).view(host.n_graphs, self.P)
        if host.graph_id is not None:
            deg = deg[torch.from_numpy(host.graph_id).long()]
        else:
            deg = deg.expand(len(graph_list), self.P)
        return deg.to(device=device, dtype=torch.float32).reshape(len(graph_list), self.P, 1, 1)

    def compute_Atx(self, x):
        """Atx[:,p] = A[0,p]^T x[:,p]  (reference :120-124), on the GPU."""
        DF.require_cuda(x)
        return DF.atx(self._operators(x.device)[0], x)

    def compute_delta(self, graph_list, y1, y2=None, device=None, _graph=None):
        """Neighbour-difference operator (reference :127-140): delta = 2*L*y1 when y2 is None."""
        DF.require_cuda(y1)
        graph = _graph if _graph is not None else BatchGraph.from_graph_list(graph_list, self.P, y1.device)
        if y2 is not None and y2 is not y1:
            # two-argument form, never used by the reference drivers: dense adjacency, off the hot path
            Bn = y1.shape[0]
            adj = torch.zeros((graph.n_graphs, self.P, self.P), dtype=y1.dtype, device=y1.device)
            ptr, idx = graph.ev_ptr.cpu(), graph.ev_idx.cpu()
            for node in range(graph.n_graphs * self.P):
                for e in idx[ptr[node]:ptr[node + 1]].tolist():
                    adj[node // self.P, node % self.P, e] += 0.5     # each neighbour appears twice in the event list
            if graph.graph_id is not None:
                adj = adj[graph.graph_id.long()]
            else:
                adj = adj.expand(Bn, self.P, self.P)
            deg = adj.sum(-1, keepdim=True).unsqueeze(-1)
            out = deg * y1 - torch.einsum("bpq,bqnk->bpnk", adj, y2)        # sum_j (y1_p - y2_j)
            out = out - (torch.einsum("bqp,bqnk->bpnk", adj, y1) - deg * y2)  # - sum_{p: q in N(p)} (y1_p - y2_q)
            return out
        y = y1.squeeze(-1).contiguous()
        zero = torch.zeros_like(y)
        hyp0 = torch.zeros((self.P, 4), dtype=y.dtype, device=y.device)
        inf = float("inf")
        _, _, d, _ = DF.step_fwd(graph, (inf, inf, inf, inf), hyp0, y, zero, zero, zero, zero,
                                 want_delta=True, want_U=False, want_graw=False)
        return d.unsqueeze(-1)

    def compute_loss(self, y_k, label):
        """mean over agents of mse(y_k[:,p], label)  (reference :142-146)."""
        return ((y_k - label.unsqueeze(1)) ** 2).mean(dim=(0, 2, 3)).sum() / self.P
