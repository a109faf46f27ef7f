"""Shared test helpers: golden-fixture loading, graph reconstruction, error metrics."""
from __future__ import annotations

import os
import sys

import networkx as nx
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "hyperparameter-gnn_unfolded-d-admm-main_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

MODEL1_CASES = ["m1_zero_P5_n64", "m1_trained_P5_n500", "m1_same_pergraph_P8_n48", "m1_trained15_P5_n51", "m1_cfg4like_P12_n40"]
MODEL3_CASES = ["m3_frozen_P5_n32"]
# the config-4-like miniature (12 agents, sparse bridged per-problem graphs) is also kept under its own name
MODEL1_EXTRA_CASES = []


def graph_from_adj(ptr, idx, P):
    """Rebuild an nx.Graph whose ``neighbors(p)`` iteration order equals the stored adjacency
    order (the reference's accumulation order depends on it)."""
    g = nx.Graph()
    g.add_nodes_from(range(P))
    for p in range(P):
        for j in idx[ptr[p]:ptr[p + 1]]:
            j = int(j)
            if j not in g._adj[p]:
                g._adj[p][j] = g._adj[j].get(p, {})
    return g


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        self.z = z
        self.name = name
        self.P, self.n, self.m, self.K, self.B = (int(z[k]) for k in ("P", "n", "m", "K", "B"))
        uniq = [graph_from_adj(z[f"adj_ptr_{i}"], z[f"adj_idx_{i}"], self.P) for i in range(int(z["n_graphs"]))]
        self.graphs = [uniq[int(i)] for i in z["graph_id"]]

    def t(self, key, dtype=None):
        if key in ("y0", "U0", "d0"):       # unfolded_DLASSO.py:49-51: randn * 1e-2 in the run's dtype
            raw = torch.from_numpy(np.asarray(self.z["noise_" + key[0]]))
            return raw.to(dtype or torch.float32) * 1e-2
        x = torch.from_numpy(np.asarray(self.z[key]))
        return x if dtype is None else x.to(dtype)

    def has(self, key):
        return key in self.z.files


def rel_l2(a, b):
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    return float(((a - b).norm() / b.norm().clamp_min(1e-300)).detach())


def random_problem(P, n, m, B, K, seed=0, graph_prob=0.5, per_sample_graphs=True, a_scale=0.1, dtype=torch.float32):
    """Seeded synthetic problem in the reference's format (set_A conditioning gnn_dlasso_utils.py:4-16,
    set_Data gnn_data.py:6-15, bridged ER graphs gnn_dlasso_progressive.py:181-191)."""
    gen = torch.Generator().manual_seed(seed)
    A = torch.zeros((1, P, m, n))
    for p in range(P):
        U, S, V = torch.svd(torch.randn((m, n), generator=gen))
        A[0, p] = U @ torch.diag(torch.clamp(S, 0.1, 10.0)) @ V.T
    A = A * a_scale
    label = 2 * torch.randn((B, n, 1), generator=gen) * (torch.rand((B, n, 1), generator=gen) <= 0.25)
    b = torch.stack([A[0, p] @ label for p in range(P)], dim=1)          # [B,P,m,1]
    graphs = []
    for i in range(B if per_sample_graphs else 1):
        g = nx.erdos_renyi_graph(P, graph_prob, seed=seed * 1000 + i)
        if not nx.is_connected(g):
            comps = list(nx.connected_components(g))
            for c in range(len(comps) - 1):
                g.add_edge(list(comps[c])[0], list(comps[c + 1])[0])
        graphs.append(g)
    if not per_sample_graphs:
        graphs = graphs * B
    y0, U0, d0 = (torch.randn((B, P, n, 1), generator=gen) * 1e-2 for _ in range(3))
    param = torch.randn((K, P, 4), generator=gen) * 0.3
    return dict(A=A.to(dtype), b=b.to(dtype), label=label.to(dtype), graphs=graphs,
                y0=y0.to(dtype), U0=U0.to(dtype), d0=d0.to(dtype), param=param.to(dtype))
