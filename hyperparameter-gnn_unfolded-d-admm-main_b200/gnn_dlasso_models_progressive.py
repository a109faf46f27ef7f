"""Drop-in replacement of the reference module ``gnn_dlasso_models_progressive.py``.

  * ``DLASSO_GNNHyp3_Progressive`` (reference :75-277): the D-ADMM recurrence (:198-232) runs in
    ``libdadmm_sm100.so`` as one differentiable kernel per iteration (``dadmm_b200.functional.Step``);
    the hypernetwork sits between iterations, so the autograd boundary is per iteration.  ``AtA y`` is
    contracted once per iteration and reused (the reference computes it twice, :158-162 and :199-203,
    and re-uploads ``AtA`` host->device each time).
  * ``GNNHypernetwork3`` (reference :9-72): the per-sample Python loop over ``torch_geometric`` calls is
    replaced by a batched dense GCN (``A_hat [B,P,P]`` bmm).  ``torch_geometric`` is not required: the
    ``GCNConv`` below keeps PyG's parameter names (``lin.weight``, ``bias``) so reference checkpoints
    load.  BatchNorm keeps the reference's per-sample semantics (statistics over the P nodes of ONE
    graph; running statistics updated sample by sample, here in closed form).  Dropout draws one mask
    per batch instead of one per sample, so train-mode outputs match the reference in distribution
    only; eval-mode outputs match it to rounding.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from dadmm_b200 import functional as DF
from dadmm_b200.graph import BatchGraph


class GCNConv(nn.Module):
    """Dense-batch graph convolution with PyG ``GCNConv`` semantics and parameter names:
    ``out = D^-1/2 (Adj + I) D^-1/2 (x W^T) + bias``; ``lin`` has no bias of its own."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lin = nn.Linear(int(in_channels), int(out_channels), bias=False)
        self.bias = nn.Parameter(torch.zeros(int(out_channels)))

    def forward(self, x, adj_hat):          # x [B,P,in], adj_hat [B,P,P]
        return torch.baddbmm(self.bias, adj_hat, self.lin(x))


def normalized_adjacency(graph_list, P, device, dtype=torch.float32):
    """A_hat [B,P,P] = D^-1/2 (Adj + I) D^-1/2 per sample, built once per distinct graph object (vectorised over the
    graphs: ``dadmm_b200.graph.normalized_adjacency_np``); ``graph_list`` may be a prebuilt ``BatchGraph``."""
    from dadmm_b200 import graph as G
    if isinstance(graph_list, BatchGraph):
        return graph_list.normalized_adjacency(dtype).to(device)
    uniq, gid, cnt, flat = G._extract(graph_list, P)
    m = torch.from_numpy(G.normalized_adjacency_np(cnt, flat, len(uniq), P)).to(device=device, dtype=dtype)
    return m.index_select(0, torch.as_tensor(gid, device=device).long()) if len(uniq) > 1 else m.expand(len(graph_list), P, P)


_EMA_CACHE = {}


def _ema_weights(B, mom, device, dtype):
    """[1,B] row of mom * (1-mom)^(B-1-b): the closed form of B sequential running-stat updates (b = 0 first)."""
    key = (B, float(mom), str(device), dtype)
    w = _EMA_CACHE.get(key)
    if w is None:
        w = (mom * (1.0 - mom) ** torch.arange(B - 1, -1, -1, device=device, dtype=torch.float64)).to(dtype).unsqueeze(0)
        if len(_EMA_CACHE) > 16:
            _EMA_CACHE.clear()
        _EMA_CACHE[key] = w
    return w


class _PerSampleBatchNorm(torch.autograd.Function):
    """Training-mode ``BatchNorm1d`` applied to every sample's [P,C] block of x [B,P,C] separately (the reference calls
    ``bn(x_b)`` sample by sample, :52-68): statistics over the P nodes of ONE graph.  Written as one autograd node with
    the textbook backward -- six kernels forward, eleven backward, where the composed element-wise formula left autograd
    some thirty small nodes per layer and iteration (the model-#3 step is bound by kernel launches, not by bytes).
    Returns (out, mean [B,1,C], biased var [B,1,C]); the statistics carry no gradient (running-stat bookkeeping only)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        ctx.set_materialize_grads(False)          # mean / var never carry a gradient: no zero tensors made for them
        var, mean = torch.var_mean(x, dim=1, unbiased=False, keepdim=True)
        rstd = torch.rsqrt(var + eps)
        xhat = (x - mean) * rstd
        ctx.save_for_backward(xhat, rstd, weight)
        ctx.mark_non_differentiable(mean, var)
        return torch.addcmul(bias, xhat, weight), mean, var

    @staticmethod
    def backward(ctx, g, _gmean, _gvar):
        if g is None:
            return None, None, None, None
        xhat, rstd, weight = ctx.saved_tensors
        gxhat = g * weight
        gw = (g * xhat).sum(dim=(0, 1)) if ctx.needs_input_grad[1] else None
        gb = g.sum(dim=(0, 1)) if ctx.needs_input_grad[2] else None
        gx = None
        if ctx.needs_input_grad[0]:
            m1 = gxhat.mean(dim=1, keepdim=True)
            m2 = (gxhat * xhat).mean(dim=1, keepdim=True)
            gx = (gxhat - m1 - xhat * m2) * rstd
        return gx, gw, gb, None


class GNNHypernetwork3(nn.Module):
    def __init__(self, P, m, hidden_dim):
        super().__init__()
        self.P, self.m = P, m
        h = int(hidden_dim)
        widths = [(self.m, h), (h, 2 * h), (2 * h, 4 * h), (4 * h, 4 * h), (4 * h, 4 * h)]
        for i, (cin, cout) in enumerate(widths, start=1):
            setattr(self, f"conv{i}", GCNConv(cin, cout))
        self.dropout = nn.Dropout(0.1)
        self.norm = nn.LayerNorm(4 * h)
        for i, (_, cout) in enumerate(widths, start=1):
            setattr(self, f"bn{i}", nn.BatchNorm1d(cout))
        for i in range(1, 6):
            nn.init.xavier_uniform_(getattr(self, f"conv{i}").lin.weight)

    @staticmethod
    def _per_sample_bn(bn, x):
        """``bn(x_b)`` for every sample b with x_b [P,C] -- x [B,P,C] -- including the sequential running-stat
        updates the reference's per-sample calls perform."""
        if not (bn.training or not bn.track_running_stats):
            return F.batch_norm(x.transpose(1, 2), bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0,
                                bn.eps).transpose(1, 2)
        Bn, Pn, _ = x.shape
        out, mean, var = _PerSampleBatchNorm.apply(x, bn.weight, bn.bias, bn.eps)
        if bn.track_running_stats:
            with torch.no_grad():
                mom = 0.1 if bn.momentum is None else bn.momentum
                w = _ema_weights(Bn, mom, x.device, x.dtype)            # [1,B] weights of the B sequential updates (cached)
                keep = (1.0 - mom) ** Bn
                # running <- keep * running + sum_b w_b stat_b, one addmm each (the running variance is the unbiased one)
                # (in place on a [1,C] view of the buffer: no temporary, no copy back)
                bn.running_mean.unsqueeze(0).addmm_(w, mean[:, 0], beta=keep)
                bn.running_var.unsqueeze(0).addmm_(w, var[:, 0], beta=keep, alpha=Pn / max(Pn - 1, 1))
        return out

    def _layer_fused(self, conv, bn, x, adj_hat, drop):
        """One layer through ``DF.GCNEpilogue``: dense product by cuBLAS, then ONE kernel for adjacency mix, bias, LeakyReLU,
        the per-graph BatchNorm and the dropout mask (CUDA, float32, P <= 64)."""
        Bn, Pn, _ = x.shape
        training_bn = bn.training or not bn.track_running_stats
        mask = None
        if drop and self.dropout.training and self.dropout.p > 0:
            keep = 1.0 - self.dropout.p
            mask = torch.empty((Bn, Pn, conv.lin.out_features), dtype=x.dtype, device=x.device).bernoulli_(keep).div_(keep)
        # (DF.linear -- the product on the library's tensor-core contraction -- was measured and lost at these sizes: every
        # call re-splits both operands into fp16 pairs, 24.6 -> 27.9 ms per graphed step at configs[1]; opt-in via tc_linear)
        if getattr(self, "tc_linear", False):
            H = DF.linear(x, conv.lin.weight)
        elif getattr(self, "split_k", True) and x.requires_grad | conv.lin.weight.requires_grad:
            H = DF.SplitKLinear.apply(x, conv.lin.weight)          # same products; the weight gradient as a split reduction
        else:
            H = conv.lin(x)
        out, mean, var = DF.GCNEpilogue.apply(H, adj_hat, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                              training_bn, bn.eps, 0.01, mask)
        if training_bn and bn.track_running_stats:
            with torch.no_grad():          # the B sequential running-stat updates of the reference's per-sample calls, in closed form
                mom = 0.1 if bn.momentum is None else bn.momentum
                w = _ema_weights(Bn, mom, x.device, x.dtype)
                keep_r = (1.0 - mom) ** Bn
                bn.running_mean.unsqueeze(0).addmm_(w, mean, beta=keep_r)
                bn.running_var.unsqueeze(0).addmm_(w, var, beta=keep_r, alpha=Pn / max(Pn - 1, 1))
        return out

    def forward(self, x, graph_list, adj_hat=None):
        """x [B,P,m,1] -> [B, P*4*hidden]  (reference :37-72)."""
        batch_size = x.shape[0]
        x = x.squeeze(-1)
        if adj_hat is None:
            adj_hat = normalized_adjacency(graph_list, self.P, x.device, x.dtype)
        fused = (getattr(self, "fused", True) and x.is_cuda and x.dtype == torch.float32 and self.P <= 64
                 and not torch.is_autocast_enabled())
        for i in range(1, 6):
            if fused:
                x = self._layer_fused(getattr(self, f"conv{i}"), getattr(self, f"bn{i}"), x, adj_hat, drop=(i < 5))
                if i == 5:
                    x = self.norm(x)
                continue
            x = F.leaky_relu(getattr(self, f"conv{i}")(x, adj_hat))
            x = self._per_sample_bn(getattr(self, f"bn{i}"), x)
            x = self.dropout(x) if i < 5 else self.norm(x)
        # the B per-sample calls of the reference bump every layer's counter B times: one fused add for the five layers
        counters = [bn.num_batches_tracked for bn in (getattr(self, f"bn{i}") for i in range(1, 6))
                    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None]
        if counters:
            with torch.no_grad():
                torch._foreach_add_(counters, batch_size)
        return x.reshape(batch_size, -1)


    def graph_conv(self, x, nx_graph):
        """One sample, the reference's per-sample entry point (:42-72): x [P,m,1] (or [P,m]) -> flat [P*4*hidden].
        ``forward`` never calls it -- the batch goes through the five layers at once; a batch of one takes the same path,
        running statistics included."""
        x = x.squeeze(-1) if x.dim() == 3 else x
        return self.forward(x.unsqueeze(0).unsqueeze(-1), [nx_graph]).reshape(-1)


class DLASSO_GNNHyp3_Progressive(nn.Module):
    def __init__(self, A, args):
        super().__init__()
        self.A = A                                   # [1,P,m,n]; plain attribute (stays where the caller put it)
        _, self.P, self.m, self.n = self.A.shape
        self.K = args.GHN_iter_num
        hidden_dim = int(args.GHyp_hidden)           # the flag is declared type=float upstream
        self.DADMM_mode = args.DADMM_mode
        self.encoder = GNNHypernetwork3(P=self.P, m=self.n * 2, hidden_dim=hidden_dim)
        layers, width = [], self.P * 4 * hidden_dim
        for out_w in (4 * hidden_dim, 2 * hidden_dim, hidden_dim):
            layers += [nn.Linear(width, out_w), nn.Dropout(0.1), nn.LayerNorm(out_w), nn.LeakyReLU()]
            width = out_w
        self.decoder = nn.Sequential(*layers)
        self.fc = nn.Linear(hidden_dim, 4 if args.DADMM_mode == 'same' else 4 * self.P)
        nn.init.xavier_uniform_(self.fc.weight, gain=0.1)
        nn.init.zeros_(self.fc.bias)
        with torch.no_grad():
            # reference :118-123 -- with the [B,4,P] view these four land on alpha of agents 0..3
            self.fc.bias[:4] = torch.tensor([-0.5, -1.0, -0.8, -1.2])[: self.fc.bias.numel()]
        self.alpha_max = torch.tensor(args.alpha_max)
        self.tau_max = torch.tensor(args.tau_max)
        self.rho_max = torch.tensor(args.rho_max)
        self.eta_max = torch.tensor(args.eta_max)
        self.contract_algo = "auto"
        self.check_finite = True
        self._ops = {}

    def _operators(self, device, dtype=None):
        return DF.solver_operators(self._ops, self.A, device, dtype or torch.get_default_dtype())

    @property
    def AtA(self):
        dev = self.A.device if self.A.is_cuda else torch.device("cuda", torch.cuda.current_device())
        return self._operators(dev)[1].unsqueeze(0).to(self.A.device)

    def hyperparameters(self, AtAy, Atb, graph_list, adj_hat=None):
        """Hypernetwork on cat([AtAy, Atb]) -> (alpha, tau, rho, eta), each [B, P|1, 1, 1]  (reference :165-196)."""
        return self._hyper_packed(AtAy, Atb, graph_list, adj_hat).unbind(dim=1)

    def _hyper_packed(self, AtAy, Atb, graph_list, adj_hat=None):
        B = AtAy.shape[0]
        h = torch.cat([AtAy, Atb], dim=2)
        h = self.fc(self._decode(self.encoder(h, graph_list, adj_hat)))
        return self._scaled(h, B)

    def _decode(self, x):
        """``self.decoder(x)`` (reference :93-106) with the Linear layers' products on the tensor-core contraction where
        their shapes take it (``DF.linear``); same modules, same parameters, same order."""
        if not (getattr(self.encoder, "tc_linear", False) and x.is_cuda and x.dtype == torch.float32):
            return self.decoder(x)
        for mod in self.decoder:
            x = DF.linear(x, mod.weight, mod.bias) if isinstance(mod, nn.Linear) else mod(x)
        return x

    def _scaled(self, h, B):
        """fc output [B, 4*(P|1)] -> [B, 4, P|1, 1, 1] = (alpha, tau, rho, eta): sigmoid, clamp, times the four maxima, tau /
        rho / eta capped at 0.9999 (reference :170-196).  The four maxima are 0-dim CPU tensors (reference attributes); the
        per-channel scale and cap live in two cached device vectors, so the four multiplies and three clamps of the
        reference are one multiply and one clamp with the same fp32 rounding and the same (closed-interval) gradient mask."""
        h = torch.clamp(torch.sigmoid(h), min=1e-4, max=0.9999)
        h = h.view(B, 4, 1 if self.DADMM_mode == 'same' else self.P, 1, 1)
        key = (str(h.device), h.dtype) + tuple(float(t) for t in (self.alpha_max, self.tau_max, self.rho_max, self.eta_max))
        if getattr(self, "_scale_key", None) != key:
            self._scale_vec = torch.tensor(key[2:], dtype=h.dtype, device=h.device).view(1, 4, 1, 1, 1)
            self._cap_vec = torch.tensor([float("inf"), 0.9999, 0.9999, 0.9999], dtype=h.dtype, device=h.device).view(1, 4, 1, 1, 1)
            self._scale_key = key
        return torch.clamp(h * self._scale_vec, max=self._cap_vec)

    def forward(self, b, graph_list, training_iterations=None, noise=None):
        """b [B,P,m,1] -> (Y [K,B,P,n,1], (alpha_k, tau_k, rho_k, eta_k) of the last iteration).
        ``noise`` (not in the reference): initial (y, U, delta) [B,P,n,1] each, replacing the three ``randn * 1e-2`` draws
        (batch-sharded runs pass their slice of the full batch's draws, ``dadmm_b200.dist.sharded_noise``)."""
        if len(b) != len(graph_list):
            raise ValueError(f"len(b)={len(b)} != len(graph_list)={len(graph_list)}")
        K = training_iterations if training_iterations is not None else self.K
        DF.require_cuda(b)
        device, B = b.device, len(b)
        A, W, Wt, At = self._operators(device)
        Atb = DF.contract(At, b.to(W.dtype).squeeze(-1), algo=self.contract_algo).unsqueeze(-1)   # [B,P,n,1]
        graph = BatchGraph.from_graph_list(graph_list, self.P, device)
        adj_hat = graph.normalized_adjacency(W.dtype)                 # cached on the BatchGraph (itself cached per graph_list)
        if noise is None:
            # torch.randn(...) * 1e-2 three times, in the reference's order (:49-51); scale applied by the generator kernel
            y0, U0, d0 = DF.initial_noise((B, self.P, self.n, 1), device)
        else:
            y0, U0, d0 = (t.to(device=device, dtype=W.dtype).reshape(B, self.P, self.n, 1) for t in noise)
        deferred = self.check_finite == "deferred"        # sticky flags, read by nonfinite_seen() only (dadmm_b200/graphs.py)
        if deferred:
            flags = self._sticky_flags(device, max(K, 1))
        else:
            flags = torch.zeros(max(K, 1), dtype=torch.int32, device=device) if self.check_finite else None
        out = self._iterate(K, W, Wt, Atb, y0, U0, d0, graph, graph_list, adj_hat, flags, guarded=False)
        if flags is not None and not deferred:
            if bool(flags.any()):
                out = self._iterate(K, W, Wt, Atb, y0, U0, d0, graph, graph_list, adj_hat, None, guarded=True)
            else:
                out[0]._dadmm_finite = out[0]._version      # no kernel saw a non-finite value (compute_loss skips its scan)
        return out

    def _sticky_flags(self, device, K):
        buf = self.__dict__.get("_flag_buf")
        if buf is None or buf.device != device or buf.numel() < K:
            buf = torch.zeros(max(K, self.K), dtype=torch.int32, device=device)
            self.__dict__["_flag_buf"] = buf
        return buf

    def nonfinite_seen(self, reset=True):
        """``check_finite = "deferred"``: has any forward pass since the last call met a non-finite value?  (one host read)"""
        buf = self.__dict__.get("_flag_buf")
        if buf is None:
            return False
        hit = bool(buf.any())
        if reset and hit:
            buf.zero_()
        return hit

    def _iterate(self, K, W, Wt, Atb, y0, U0, d0, graph, graph_list, adj_hat, flags, guarded):
        bad = lambda t: bool(torch.isnan(t).any() or torch.isinf(t).any())
        y, U, d = y0.squeeze(-1), U0.squeeze(-1), d0.squeeze(-1)
        Atb3 = Atb.squeeze(-1)
        clamps = DF.clamps_model3()
        Y, hyp = [], None
        packed_path = type(self).hyperparameters is DLASSO_GNNHyp3_Progressive.hyperparameters
        for k in range(K):
            if guarded:
                if bad(y):
                    print(f"Warning: NaN/Inf detected in y_k at iteration {k}, resetting...")
                    y = torch.zeros_like(y)
                if bad(U):
                    print(f"Warning: NaN/Inf detected in U_k at iteration {k}, resetting...")
                    U = torch.zeros_like(U)
            AtAy = DF.Contract.apply(y, W, Wt, self.contract_algo)            # [B,P,n]
            if packed_path:
                # the four outputs are views of one [B,4,P|1,1,1] tensor, which is also the step kernel's [B,4,P] layout
                packed = self._hyper_packed(AtAy.unsqueeze(-1), Atb, graph_list, adj_hat)
                hyp = packed.unbind(dim=1)
                hyp_s = packed.reshape(len(y), 4, -1).expand(len(y), 4, self.P).contiguous()
            else:                       # ``hyperparameters`` overridden by a subclass: honour it
                hyp = self.hyperparameters(AtAy.unsqueeze(-1), Atb, graph_list, adj_hat)
                hyp_s = torch.stack([h.reshape(len(y), -1).expand(len(y), self.P) for h in hyp], dim=1).contiguous()
            flag = torch.zeros(1, dtype=torch.int32, device=y.device) if guarded else (flags[k:k + 1] if flags is not None else None)
            y_n, U_n, d_n = DF.Step.apply(y, U, d, AtAy, Atb3, hyp_s, graph, clamps, flag)
            if guarded:
                if int(flag) & 4:
                    print(f"Warning: NaN/Inf in gradient at iteration {k}, skipping update...")
                    y_n = torch.clamp(y, -clamps[1], clamps[1])
                    zero = torch.zeros_like(y_n)
                    _, _, d_n, _ = DF.step_fwd(graph, (float("inf"),) * 2 + (clamps[2], float("inf")),
                                               torch.zeros((self.P, 4), dtype=y.dtype, device=y.device), y_n.detach(), zero,
                                               zero, zero, zero, want_U=False, want_graw=False)
                    U_n = torch.clamp(U + d_n * hyp_s[:, 3].unsqueeze(-1), -clamps[3], clamps[3])
                if bad(y_n):
                    print(f"Warning: NaN/Inf in y_next at iteration {k}, using previous value...")
                    y_n = y
            y, U, d = y_n, U_n, d_n
            Y.append(y)
        return torch.stack(Y).unsqueeze(-1), hyp

    # ------------------------------------------------------------------ reference helper API
    def compute_sum_neighbors(self, graph_list):
        params = list(self.parameters())
        device = params[0].device if params else torch.device("cpu")
        host = BatchGraph.build_host(graph_list, self.P)
        deg = torch.from_numpy(host.deg).view(host.n_graphs, self.P)
        deg = deg[torch.from_numpy(host.graph_id).long()] if host.graph_id is not None else deg.expand(len(graph_list), self.P)
        return deg.to(device=device, dtype=torch.float32).reshape(len(graph_list), self.P, 1, 1)

    def compute_Atx(self, x):
        DF.require_cuda(x)
        return DF.atx(self._operators(x.device)[0], x)

    def compute_delta(self, graph_list, y1, y2=None):
        if y2 is not None and y2 is not y1:
            raise NotImplementedError("two-argument compute_delta is not used on the D-ADMM path")
        DF.require_cuda(y1)
        graph = BatchGraph.from_graph_list(graph_list, self.P, y1.device)
        y = y1.squeeze(-1).contiguous()
        zero = torch.zeros_like(y)
        inf = float("inf")
        _, _, d, _ = DF.step_fwd(graph, (inf, inf, inf, inf), torch.zeros((self.P, 4), dtype=y.dtype, device=y.device),
                                 y, zero, zero, zero, zero, want_U=False, want_graw=False)
        return d.unsqueeze(-1)
