"""``graph_list`` (one ``networkx.Graph`` per problem) -> device CSR consumed by the kernels.

The reference walks the graphs in Python on every iteration (``compute_delta``,
unfolded_DLASSO.py:127-140; ``compute_sum_neighbors``, :111-118).  Here each DISTINCT graph object
of a batch (the drivers pass ``[graph]*B`` or B fresh graphs) is converted once into an
*event-ordered* neighbour list -- the neighbour ids in the exact order the reference accumulates
``y_q - y_e`` into ``delta[q]`` -- so the kernels reproduce ``delta = 2*L*y`` with the reference's
rounding, and ``deg[p] = len(list(graph.neighbors(p)))``.

``sample_erdos_renyi`` makes such a batch without ``networkx``: the driver's per-problem generation loop
(gnn_dlasso_progressive.py:181-191) as tensor passes on the device.
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


def _ordered_adjacency(graph, P: int):
    """Neighbour ids of nodes 0..P-1 in ``graph.neighbors`` order, flattened, plus the per-node counts.  networkx keeps
    that order in ``_adj`` (dict of dicts); other graph-likes go through ``neighbors()``."""
    adj = getattr(graph, "_adj", None)
    if isinstance(adj, dict):
        try:
            rows = [adj[p] for p in range(P)]
        except KeyError as e:
            raise ValueError(f"graph has no node {e.args[0]!r}: nodes must be 0..{P - 1}") from None
    else:
        rows = [list(graph.neighbors(p)) for p in range(P)]
    return rows


def _extract(graph_list: Sequence, P: int):
    """The only per-graph Python work of the ingestion: de-duplicate the graph objects of the batch and read their
    neighbour order.  Returns (unique graphs, graph_id[B] int32, cnt[G*P] int64, flat[sum cnt] int64) with ``flat`` the
    neighbour ids in visit order (graph, p = 0..P-1, ``graph.neighbors(p)`` order)."""
    from itertools import chain
    uniq, index, gid = [], {}, np.empty(len(graph_list), np.int32)
    for b, g in enumerate(graph_list):
        k = id(g)
        if k not in index:
            index[k] = len(uniq)
            uniq.append(g)
        gid[b] = index[k]
    cnt_l, flat_l = [], []
    for g in uniq:
        rows = _ordered_adjacency(g, P)
        cnt_l.extend(map(len, rows))
        flat_l.extend(chain.from_iterable(rows))
    cnt = np.asarray(cnt_l, np.int64)
    try:
        flat = np.asarray(flat_l, np.int64)
    except (TypeError, ValueError, OverflowError):
        raise ValueError(f"graph nodes must be the integers 0..{P - 1}") from None
    if flat.size and (flat.min() < 0 or flat.max() >= P):
        bad = flat_l[int(np.argmax((flat < 0) | (flat >= P)))]
        raise ValueError(f"graph node {bad!r} outside 0..{P - 1}")
    return uniq, gid, cnt, flat


class HostGraph:
    """Host-side (numpy int32) CSR arrays of the distinct graphs of one batch.

    Only the neighbour order is read graph by graph in Python (``_extract``); event lists, plain adjacency and degrees
    of all graphs are then built with a handful of vectorised passes -- a stable sort of the (owner, value) event
    stream reproduces ``event_lists`` exactly.  4096 fresh 50-node graphs on the GPU box's host: 0.17 s, or 0.14 s when
    the passes run on the GPU (``BatchGraph.from_graph_list`` on a CUDA device) -- what remains is the Python read of
    the neighbour order; the per-graph construction this replaces took ~5x longer."""

    def __init__(self, graph_list: Sequence, P: int):
        uniq, gid, cnt, flat = _extract(graph_list, P)
        G = len(uniq)
        node = np.repeat(np.arange(G * P, dtype=np.int64), cnt)  # global id (g*P + p) of the visiting node
        gbase = node - node % P
        dst = gbase + flat                                       # global id of the neighbour
        # event stream of unfolded_DLASSO.py:132-139: visit (p, j) appends j to ev[p] and p to ev[j]
        V = flat.size
        owner = np.empty(2 * V, np.int64)
        val = np.empty(2 * V, np.int64)
        owner[0::2], val[0::2] = node, flat
        owner[1::2], val[1::2] = dst, node - gbase
        order = np.argsort(owner, kind="stable")
        ev_idx = val[order]
        ev_cnt = np.bincount(owner, minlength=G * P)
        keep = dst != node                                       # plain adjacency: each neighbour once, self-loops dropped
        adj_idx = flat[keep]
        adj_cnt = np.bincount(node[keep], minlength=G * P)
        ptr = lambda c: np.concatenate(([0], np.cumsum(c))).astype(np.int32)
        arr = lambda x: (x if x.size else np.zeros(1, np.int64)).astype(np.int32)   # device pointers stay non-null for edgeless graphs
        self.ev_ptr, self.ev_idx = ptr(ev_cnt), arr(ev_idx)
        self.adj_ptr, self.adj_idx = ptr(adj_cnt), arr(adj_idx)
        self.deg = cnt.astype(np.int32) if cnt.size else np.zeros(1, np.int32)
        self.graph_id = gid if G > 1 else None
        self.n_graphs, self.P, self.B = G, P, len(graph_list)
        per_graph = lambda c: int(c.reshape(G, P).sum(axis=1).max()) if G and P else 0
        self.max_events, self.max_adj = per_graph(ev_cnt), per_graph(adj_cnt)
        self.unique_graphs = uniq
        self.nb_cnt, self.nb_flat = cnt, flat          # neighbour lists as read (normalized_adjacency)


class _DeviceCSR:
    """Same fields as ``HostGraph``, built on a CUDA device with torch ops from the extracted neighbour order (the
    event stream's stable sort, the counts and prefix sums run on the GPU)."""

    def __init__(self, graph_list: Sequence, P: int, device):
        uniq, gid, cnt_h, flat_h = _extract(graph_list, P)
        dev = torch.device(device)
        self._build(len(uniq), P, len(graph_list), gid, torch.from_numpy(cnt_h).to(dev), torch.from_numpy(flat_h).to(dev), dev)
        self.unique_graphs = uniq
        self.nb_cnt, self.nb_flat = cnt_h, flat_h

    @classmethod
    def from_lists(cls, G: int, P: int, cnt: torch.Tensor, flat: torch.Tensor) -> "_DeviceCSR":
        """G distinct graphs, one per problem, from neighbour lists already on the device (``sample_erdos_renyi``):
        cnt[G*P] list lengths, flat = neighbour ids in visit order (both int64)."""
        self = object.__new__(cls)
        self._build(G, P, G, np.arange(G, dtype=np.int32), cnt, flat, cnt.device)
        self.unique_graphs = None
        self.nb_cnt, self.nb_flat = cnt, flat          # read back lazily (BatchGraph.normalized_adjacency)
        return self

    def _build(self, G, P, B, gid, cnt, flat, dev):
        node = torch.repeat_interleave(torch.arange(G * P, device=dev), cnt)
        gbase = node - node % P
        dst = gbase + flat
        owner = torch.stack((node, dst), dim=1).reshape(-1)
        val = torch.stack((flat, node - gbase), dim=1).reshape(-1)
        order = torch.sort(owner, stable=True).indices
        ev_cnt = torch.bincount(owner, minlength=G * P)
        keep = dst != node
        adj_cnt = torch.bincount(node[keep], minlength=G * P)
        zero = torch.zeros(1, dtype=torch.int64, device=dev)
        ptr = lambda c: torch.cat((zero, torch.cumsum(c, 0))).to(torch.int32)
        arr = lambda x: (x if x.numel() else zero).to(torch.int32)
        self.ev_ptr, self.ev_idx = ptr(ev_cnt), arr(val[order])
        self.adj_ptr, self.adj_idx = ptr(adj_cnt), arr(flat[keep])
        self.deg = cnt.to(torch.int32) if cnt.numel() else torch.zeros(1, dtype=torch.int32, device=dev)
        self.graph_id = gid if G > 1 else None
        self.n_graphs, self.P, self.B = G, P, B
        if G and P:
            mx = torch.stack((ev_cnt.view(G, P).sum(1).max(), adj_cnt.view(G, P).sum(1).max())).tolist()   # one sync
            self.max_events, self.max_adj = int(mx[0]), int(mx[1])
        else:
            self.max_events = self.max_adj = 0


def normalized_adjacency_np(cnt, flat, n_graphs: int, P: int):
    """A_hat [G,P,P] = D^-1/2 (Adj + I) D^-1/2 in float64 for all distinct graphs at once, from their neighbour lists
    (``_extract``: cnt[G*P] list lengths, flat = neighbour ids).  Same arithmetic, element by element, as the per-graph
    construction it replaces (identity first, off-diagonal ones, column sums ** -0.5, row scale then column scale) --
    1024 fresh 5-node graphs took 85 ms per forward pass that way, more than the whole model-#3 training step."""
    a = np.zeros((n_graphs, P, P), np.float64)
    if n_graphs and P:
        node = np.repeat(np.arange(n_graphs * P, dtype=np.int64), cnt)
        g, u, v = node // P, node % P, np.asarray(flat, np.int64)
        off = u != v
        a[g[off], u[off], v[off]] = 1.0
        a[g[off], v[off], u[off]] = 1.0
        idx = np.arange(P)
        a[:, idx, idx] = 1.0
    d = a.sum(axis=1) ** -0.5
    return d[:, :, None] * a * d[:, None, :]


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
        self._nb = (getattr(host, "nb_cnt", None), getattr(host, "nb_flat", None))
        self._adj_hat = {}
        self.c = _lib.Graph(self.n_graphs, self.P, self.ev_ptr.data_ptr(), self.ev_idx.data_ptr(), self.deg.data_ptr(),
                            self.graph_id.data_ptr() if self.graph_id is not None else None,
                            self.adj_ptr.data_ptr(), self.adj_idx.data_ptr(), self.max_events, self.max_adj)

    @property
    def device(self):
        return self.ev_ptr.device

    @staticmethod
    def build_host(graph_list: Sequence, P: int) -> HostGraph:
        return HostGraph(graph_list, P)

    def __len__(self):
        return self.B

    def normalized_adjacency(self, dtype=torch.float32) -> torch.Tensor:
        """[B,P,P] GCN propagation matrices of the batch's problems (``normalized_adjacency_np`` per distinct graph,
        gathered by ``graph_id``), on this graph's device; built once per BatchGraph and dtype."""
        hit = self._adj_hat.get(dtype)
        if hit is None:
            cnt, flat = self._nb
            if cnt is None:
                raise ValueError("this BatchGraph carries no neighbour lists")
            if isinstance(cnt, torch.Tensor):          # sampled on the device: one read-back, kept
                cnt, flat = cnt.cpu().numpy(), flat.cpu().numpy()
                self._nb = (cnt, flat)
            m = torch.from_numpy(normalized_adjacency_np(cnt, flat, self.n_graphs, self.P)).to(device=self.device, dtype=dtype)
            hit = m.index_select(0, self.graph_id.long()) if self.graph_id is not None else m.expand(self.B, self.P, self.P)
            self._adj_hat[dtype] = hit
        return hit

    def to_networkx(self, problems=None) -> list:
        """``networkx.Graph`` objects of the given problems (default: all), edges inserted so that every node's
        ``neighbors()`` order equals the stored list order -- ``BatchGraph.from_graph_list`` of the result reproduces
        this graph's arrays.  Needs the neighbour lists (any graph built by this module has them)."""
        import networkx as nx
        cnt, flat = self._nb
        if cnt is None:
            raise ValueError("this BatchGraph carries no neighbour lists")
        if isinstance(cnt, torch.Tensor):
            cnt, flat = cnt.cpu().numpy(), flat.cpu().numpy()
            self._nb = (cnt, flat)
        P = self.P
        ptr = np.concatenate(([0], np.cumsum(cnt)))
        gid = self.graph_id.cpu().numpy() if self.graph_id is not None else np.zeros(self.B, np.int64)
        built, out = {}, []
        for b in (range(self.B) if problems is None else problems):
            g = int(gid[b])
            if g not in built:
                G = nx.Graph()
                G.add_nodes_from(range(P))
                rows = [flat[ptr[g * P + u]:ptr[g * P + u + 1]].tolist() for u in range(P)]
                # insertion order that yields these lists: repeatedly take the earliest pending edge whose two ends
                # both have it at the head of their remaining list (always exists for lists networkx could hold)
                pos = [0] * P
                remaining = sum(len(r) for r in rows)
                while remaining:
                    progressed = False
                    for u in range(P):
                        while pos[u] < len(rows[u]):
                            v = rows[u][pos[u]]
                            if v == u:
                                G.add_edge(u, u); pos[u] += 1; remaining -= 1; progressed = True
                            elif pos[v] < len(rows[v]) and rows[v][pos[v]] == u:
                                G.add_edge(u, v); pos[u] += 1; pos[v] += 1; remaining -= 2; progressed = True
                            else:
                                break
                    if not progressed:
                        raise ValueError("neighbour lists are not the adjacency order of any undirected graph")
                built[g] = G
            out.append(built[g])
        return out

    @classmethod
    def from_graph_list(cls, graph_list, P: int, device) -> "BatchGraph":
        """``graph_list``: the reference's list of ``networkx.Graph`` (one per problem), or a ``BatchGraph`` built
        earlier (returned as is: lets a data pipeline convert ahead of the step and reuse graphs across batches)."""
        if isinstance(graph_list, BatchGraph):
            if graph_list.P != P or str(graph_list.device) != str(torch.device(device)):
                raise ValueError("BatchGraph was built for another P / device")
            return graph_list
        key = (tuple(id(g) for g in graph_list), P, str(device))
        hit = _cache.get(key)
        if hit is not None:
            _cache.move_to_end(key)
            return hit
        on_gpu = torch.device(device).type == "cuda"
        csr = _DeviceCSR(graph_list, P, device) if on_gpu else HostGraph(graph_list, P)
        bg = cls(csr, device, keepalive=list(graph_list))   # strong refs: ids stay unique while cached
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
        out._adj_hat = {}
        if getattr(self, "n_bridges", None) is not None and self.graph_id is not None:
            out.n_bridges = self.n_bridges[out.graph_id.long()]
        out.c = _lib.Graph(self.n_graphs, self.P, self.ev_ptr.data_ptr(), self.ev_idx.data_ptr(), self.deg.data_ptr(),
                           out.graph_id.data_ptr() if out.graph_id is not None else None,
                           self.adj_ptr.data_ptr(), self.adj_idx.data_ptr(), self.max_events, self.max_adj)
        return out


def _erdos_renyi_lists(B: int, P: int, p: float, device, generator=None, connect: bool = True):
    """Neighbour lists of B independent G(P, p) graphs, drawn and (optionally) bridged with tensor passes on ``device``.

    Returns (cnt[B*P] int64, flat int64, n_bridges[B] int64).  List order is the one ``networkx`` would hold had each
    graph been built the way the reference's driver builds it (gnn_dlasso_progressive.py:181-191): the G(P, p) edges
    inserted in (u, v), u < v, lexicographic order -- every node's neighbours ascending -- then one bridging edge
    between consecutive connected components, appended after them."""
    dev = torch.device(device)
    idx = torch.arange(P, device=dev)
    upper = idx.view(1, P, 1) < idx.view(1, 1, P)
    er = (torch.rand((B, P, P), device=dev, generator=generator) < p) & upper
    er = er | er.transpose(1, 2)                                           # [B,P,P] symmetric, no self-loops
    bridge = torch.zeros_like(er)
    n_bridges = torch.zeros(B, dtype=torch.int64, device=dev)
    if connect and P > 1:
        # connected components: propagate the smallest node id over the edges until nothing changes (a component's
        # label is its smallest node -- the order nx.connected_components lists them in)
        label = idx.expand(B, P).contiguous()
        big = torch.full((), P, dtype=label.dtype, device=dev)
        for _ in range(P):
            nb_min = torch.where(er, label.unsqueeze(1), big).amin(dim=2)
            new = torch.minimum(label, nb_min)
            new = torch.gather(new, 1, new)                                # pointer jumping: label of my label
            if torch.equal(new, label):
                break
            label = new
        rep = label == idx                                                 # component representatives, ascending in p
        # next representative after p (P when none): reverse running minimum of the representatives' ids
        cand = torch.where(rep, idx, big).flip(1).cummin(dim=1).values.flip(1)            # min id of a rep at >= p
        nxt = torch.cat((cand[:, 1:], big.expand(B, 1)), dim=1)                           # ... at > p
        has = rep & (nxt < P)
        b_i, u_i = has.nonzero(as_tuple=True)
        v_i = nxt[b_i, u_i]
        bridge[b_i, u_i, v_i] = True
        bridge[b_i, v_i, u_i] = True
        n_bridges = has.sum(dim=1)
    both = torch.cat((er, bridge), dim=2)                                  # [B,P,2P]: per node, G(P,p) edges then bridges
    cnt = both.sum(dim=2).reshape(-1)
    flat = both.nonzero(as_tuple=True)[2] % P                              # row-major = (graph, node, list position)
    return cnt, flat, n_bridges


def sample_erdos_renyi(B: int, P: int, p: float, device, generator=None, connect: bool = True) -> "BatchGraph":
    """A batch of B fresh G(P, p) graphs, made connected the way the reference's driver does it, as a ``BatchGraph``
    -- sampling, component search, bridging and the CSR build all run as tensor passes on ``device``, so a batch of
    4096 50-agent problems does not spend seconds in ``networkx`` before every step (3 s for the generation loop of
    gnn_dlasso_progressive.py:181-191 on this container's CPU, against a 0.1 s training step).  Pass the result wherever
    the modules take ``graph_list``.

    Same distribution as the driver's loop, not the same random stream (``networkx`` draws from Python's ``random``);
    one documented difference in the bridging: the driver joins ``list(component)[0]`` of consecutive components --
    whichever element the Python set happens to yield first -- this joins their smallest nodes.
    ``BatchGraph.to_networkx`` rebuilds the ``networkx`` objects (same insertion order) when something else needs them."""
    if not (0.0 <= p <= 1.0) or B < 1 or P < 1:
        raise ValueError("sample_erdos_renyi: need B >= 1, P >= 1, 0 <= p <= 1")
    cnt, flat, n_bridges = _erdos_renyi_lists(B, P, p, device, generator, connect)
    bg = BatchGraph(_DeviceCSR.from_lists(B, P, cnt, flat), device)
    bg.n_bridges = n_bridges
    return bg


_CACHE_MAX = 8
_cache: "OrderedDict[tuple, BatchGraph]" = OrderedDict()
