"""``graph_list`` (one ``networkx.Graph`` per problem) -> device CSR consumed by the kernels.

The reference walks the graphs in Python on every iteration (``compute_delta``,
unfolded_DLASSO.py:127-140; ``compute_sum_neighbors``, :111-118).  Here each DISTINCT graph object
of a batch (the drivers pass ``[graph]*B`` or B fresh graphs) is converted once into an
*event-ordered* neighbour list -- the neighbour ids in the exact order the reference accumulates
``y_q - y_e`` into ``delta[q]`` -- so the kernels reproduce ``delta = 2*L*y`` with the reference's
rounding, and ``deg[p] = len(list(graph.neighbors(p)))``.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import List, Sequence

import numpy as np
import torch

from . import _lib


def event_lists(graph, P: int) -> List[List[int]]:
    """Accumulation order of unfolded_DLASSO.py:132-139: the outer loop visits p = 0..P-1 and for
    j in N(p) adds (y_p - y_j) to delta[p] and subtracts it from delta[j] (== adds y_j - y_p)."""
    ev: List[List[int]] = [[] for _ in range(P)]
    for p in range(P):
        for j in graph.neighbors(p):
            if not (0 <= j < P):
                raise ValueError(f"graph node {j!r} outside 0..{P - 1}")
            ev[p].append(j)
            ev[j].append(p)
    return ev


class HostGraph:
    """Host-side (numpy int32) CSR arrays of the distinct graphs of one batch."""

    def __init__(self, graph_list: Sequence, P: int):
        uniq, index, gid = [], {}, np.empty(len(graph_list), np.int32)
        for b, g in enumerate(graph_list):
            k = id(g)
            if k not in index:
                index[k] = len(uniq)
                uniq.append(g)
            gid[b] = index[k]
        ev_ptr, ev_idx, adj_ptr, adj_idx, deg = [0], [], [0], [], []
        max_events = max_adj = 0
        for g in uniq:
            ev0, adj0 = len(ev_idx), len(adj_idx)
            for p, ev in enumerate(event_lists(g, P)):
                ev_idx.extend(ev)
                ev_ptr.append(len(ev_idx))
            for p in range(P):
                nb = list(g.neighbors(p))
                deg.append(len(nb))
                adj_idx.extend(j for j in nb if j != p)
                adj_ptr.append(len(adj_idx))
            max_events = max(max_events, len(ev_idx) - ev0)
            max_adj = max(max_adj, len(adj_idx) - adj0)
        arr = lambda x: np.asarray(x if x else [0], np.int32)    # keep device pointers non-null for edgeless graphs
        self.ev_ptr, self.ev_idx, self.adj_ptr, self.adj_idx, self.deg = arr(ev_ptr), arr(ev_idx), arr(adj_ptr), arr(adj_idx), arr(deg)
        self.graph_id = gid if len(uniq) > 1 else None
        self.n_graphs, self.P, self.B = len(uniq), P, len(graph_list)
        self.max_events, self.max_adj = max_events, max_adj
        self.unique_graphs = uniq


class BatchGraph:
    """Device-resident CSR of the distinct graphs of one batch + per-problem graph index."""

    def __init__(self, host: HostGraph, device, keepalive=None, graph_id=None, B=None):
        dev = torch.device(device)
        t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(a).to(dev)
        self.ev_ptr, self.ev_idx, self.deg = t(host.ev_ptr), t(host.ev_idx), t(host.deg)
        self.adj_ptr, self.adj_idx = t(host.adj_ptr), t(host.adj_idx)
        gid = host.graph_id if graph_id is None else graph_id
        self.graph_id = t(gid) if gid is not None else None
        self.n_graphs, self.P, self.B = host.n_graphs, host.P, (host.B if B is None else B)
        self.max_events, self.max_adj = host.max_events, host.max_adj
        self._keepalive = keepalive
        self.c = _lib.Graph(self.n_graphs, self.P, self.ev_ptr.data_ptr(), self.ev_idx.data_ptr(), self.deg.data_ptr(),
                            self.graph_id.data_ptr() if self.graph_id is not None else None,
                            self.adj_ptr.data_ptr(), self.adj_idx.data_ptr(), self.max_events, self.max_adj)

    @property
    def device(self):
        return self.ev_ptr.device

    @staticmethod
    def build_host(graph_list: Sequence, P: int) -> HostGraph:
        return HostGraph(graph_list, P)

    @classmethod
    def from_graph_list(cls, graph_list: Sequence, P: int, device) -> "BatchGraph":
        key = (tuple(id(g) for g in graph_list), P, str(device))
        hit = _cache.get(key)
        if hit is not None:
            _cache.move_to_end(key)
            return hit
        bg = cls(HostGraph(graph_list, P), device, keepalive=list(graph_list))   # strong refs: ids stay unique while cached
        _cache[key] = bg
        while len(_cache) > _CACHE_MAX:
            _cache.popitem(last=False)
        return bg

    def shard(self, lo: int, hi: int) -> "BatchGraph":
        """Problems [lo, hi) of this batch (multi-GPU batch sharding); shares the CSR arrays."""
        out = object.__new__(BatchGraph)
        out.__dict__.update(self.__dict__)
        out.graph_id = self.graph_id[lo:hi].contiguous() if self.graph_id is not None else None
        out.B = hi - lo
        out.c = _lib.Graph(self.n_graphs, self.P, self.ev_ptr.data_ptr(), self.ev_idx.data_ptr(), self.deg.data_ptr(),
                           out.graph_id.data_ptr() if out.graph_id is not None else None,
                           self.adj_ptr.data_ptr(), self.adj_idx.data_ptr(), self.max_events, self.max_adj)
        return out


_CACHE_MAX = 8
_cache: "OrderedDict[tuple, BatchGraph]" = OrderedDict()
