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


class BatchGraph:
    """Device-resident CSR of the distinct graphs of one batch + per-problem graph index."""

    def __init__(self, ev_ptr, ev_idx, deg, graph_id, n_graphs, P, B, keepalive=None):
        self.ev_ptr, self.ev_idx, self.deg, self.graph_id = ev_ptr, ev_idx, deg, graph_id
        self.n_graphs, self.P, self.B = n_graphs, P, B
        self._keepalive = keepalive
        self.c = _lib.Graph(n_graphs, P, ev_ptr.data_ptr(), ev_idx.data_ptr(), deg.data_ptr(),
                            graph_id.data_ptr() if graph_id is not None else None)

    @property
    def device(self):
        return self.ev_ptr.device

    @staticmethod
    def build_host(graph_list: Sequence, P: int):
        """Host-side arrays (numpy int32): ev_ptr [G*P+1], ev_idx, deg [G*P], graph_id [B] (or None)."""
        uniq, index, gid = [], {}, np.empty(len(graph_list), np.int32)
        for b, g in enumerate(graph_list):
            k = id(g)
            if k not in index:
                index[k] = len(uniq)
                uniq.append(g)
            gid[b] = index[k]
        ptr, idx, deg = [0], [], []
        for g in uniq:
            for p, ev in enumerate(event_lists(g, P)):
                idx.extend(ev)
                ptr.append(len(idx))
            deg.extend(len(list(g.neighbors(p))) for p in range(P))
        if not idx:
            idx = [0]          # keep the device pointer non-null for edgeless graphs
        return (np.asarray(ptr, np.int32), np.asarray(idx, np.int32), np.asarray(deg, np.int32),
                (gid if len(uniq) > 1 else None), len(uniq))

    @classmethod
    def from_graph_list(cls, graph_list: Sequence, P: int, device) -> "BatchGraph":
        key = (tuple(id(g) for g in graph_list), P, str(device))
        hit = _cache.get(key)
        if hit is not None:
            _cache.move_to_end(key)
            return hit
        ptr, idx, deg, gid, G = cls.build_host(graph_list, P)
        dev = torch.device(device)
        t = lambda a: torch.from_numpy(a).to(dev)
        bg = cls(t(ptr), t(idx), t(deg), t(gid) if gid is not None else None, G, P, len(graph_list),
                 keepalive=list(graph_list))   # strong refs: ids stay unique while cached
        _cache[key] = bg
        while len(_cache) > _CACHE_MAX:
            _cache.popitem(last=False)
        return bg

    def shard(self, lo: int, hi: int) -> "BatchGraph":
        """Problems [lo, hi) of this batch (multi-GPU batch sharding)."""
        gid = self.graph_id[lo:hi].contiguous() if self.graph_id is not None else None
        return BatchGraph(self.ev_ptr, self.ev_idx, self.deg, gid, self.n_graphs, self.P, hi - lo, self._keepalive)


_CACHE_MAX = 8
_cache: "OrderedDict[tuple, BatchGraph]" = OrderedDict()
