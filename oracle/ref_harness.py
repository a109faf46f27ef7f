"""Import the UNMODIFIED reference modules from ``/root/reference`` (build container only).

TEST INFRASTRUCTURE.  Used by ``oracle/make_golden.py`` (to mint ``tests/golden/*.npz``) and by
the ``-m "not gpu"`` tests that are skipped when ``/root/reference`` is absent (the GPU box).
Nothing here is imported by the product path.

The reference imports ``torch_geometric`` at module top (unfolded_DLASSO.py:4-5,
gnn_dlasso_models_progressive.py:4-5); the package is un-vendored and unpinned
(requirements.txt:11) and absent from this image, so stub modules are injected into
``sys.modules`` before the import.  ``DLASSO_unfolded`` never touches the stub.  For model #3
the stub ``GCNConv`` follows the documented PyG semantics (``lin`` without bias, separate
zero-init ``bias``, self-loops, symmetric normalisation, sum aggregation) -- parity of the
hypernetwork against the real PyG is therefore UNPINNED; only the D-ADMM recurrence is pinned.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("DADMM_REFERENCE_ROOT", "/root/reference")
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")     # oracle/build_ref.py (travels to the GPU box)


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "unfolded_DLASSO.py"))


def reference_root(name: str = "unfolded_DLASSO") -> str:
    """Directory ``name``.py is imported from: the reference checkout where it exists (build container), else the
    unmodified copies staged under ``oracle/_ref`` (digests checked against the manifest).  '' when neither has it."""
    if os.path.isfile(os.path.join(REFERENCE_ROOT, name + ".py")):
        return REFERENCE_ROOT
    from . import build_ref
    if build_ref.staged() and os.path.isfile(os.path.join(STAGED_ROOT, name + ".py")):
        return STAGED_ROOT
    return ""


class _StubGCNConv(nn.Module):
    """PyG-semantics GCNConv: out = D^-1/2 (Adj + I) D^-1/2 (x W^T) + bias."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lin = nn.Linear(int(in_channels), int(out_channels), bias=False)
        self.bias = nn.Parameter(torch.zeros(int(out_channels)))

    def forward(self, x, edge_index):
        N = x.shape[0]
        loops = torch.arange(N, device=x.device)
        row = torch.cat([edge_index[0], loops])
        col = torch.cat([edge_index[1], loops])
        w = torch.ones(row.shape[0], dtype=x.dtype, device=x.device)
        deg = torch.zeros(N, dtype=x.dtype, device=x.device).index_add_(0, col, w)
        dis = deg.pow(-0.5)
        dis[torch.isinf(dis)] = 0
        norm = dis[row] * w * dis[col]
        h = self.lin(x)
        out = torch.zeros_like(h).index_add_(0, col, h[row] * norm.unsqueeze(-1))
        return out + self.bias


def _stub_from_networkx(G):
    edges = []
    for u, v in G.edges():
        edges.append((u, v))
        if u != v:
            edges.append((v, u))
    edges.sort()
    ei = torch.tensor(edges, dtype=torch.long).t().contiguous() if edges else torch.zeros((2, 0), dtype=torch.long)
    return types.SimpleNamespace(edge_index=ei)


def _install_stubs():
    if "torch_geometric" in sys.modules:
        return
    tg = types.ModuleType("torch_geometric")
    tgnn = types.ModuleType("torch_geometric.nn")
    tgu = types.ModuleType("torch_geometric.utils")
    tgnn.GCNConv = _StubGCNConv
    tgnn.global_mean_pool = lambda x, batch=None: x.mean(dim=0, keepdim=True)
    tgu.from_networkx = _stub_from_networkx
    tg.nn, tg.utils = tgnn, tgu
    sys.modules.update({"torch_geometric": tg, "torch_geometric.nn": tgnn, "torch_geometric.utils": tgu})


_ref_cache = {}


def load(name: str):
    """Import reference module ``name`` (e.g. 'unfolded_DLASSO') under a private alias so that it
    cannot shadow / be shadowed by the same-named drop-in modules of the product package."""
    if name in _ref_cache:
        return _ref_cache[name]
    root = reference_root(name)
    if not root:
        raise RuntimeError(f"reference module {name}.py found neither at {REFERENCE_ROOT} nor staged under {STAGED_ROOT}")
    _install_stubs()
    spec = importlib.util.spec_from_file_location(f"_dadmm_ref_{name}", os.path.join(root, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _ref_cache[name] = mod
    return mod


@contextlib.contextmanager
def default_dtype(dtype):
    """Run the reference in fp64: ``torch.set_default_dtype`` + ``torch.randn`` patched to draw in
    fp32 and cast, so that the initial noise is the same numbers in both precisions."""
    old = torch.get_default_dtype()
    orig_randn = torch.randn
    if dtype != torch.float32:
        def randn32(*a, **kw):
            kw.setdefault("dtype", torch.float32)
            return orig_randn(*a, **kw).to(dtype)
        torch.randn = randn32
    torch.set_default_dtype(dtype)
    try:
        yield
    finally:
        torch.set_default_dtype(old)
        torch.randn = orig_randn
