"""GPU run of the batched Erdos-Renyi sampler (dadmm_b200.graph.sample_erdos_renyi; its arithmetic is pinned by the CPU
tests in test_cpu_host.py -- the same tensor passes): sampled on the device, it must equal the ingestion of its own
networkx graphs, and the solver must give the same bits for the sampled BatchGraph and for that graph_list."""
import pytest
import torch

from test_gpu_chain import DEV, _module_case

pytestmark = pytest.mark.gpu


def test_sampler_on_device_feeds_the_solver():
    import networkx as nx
    import gnn_dlasso_utils
    from dadmm_b200 import graph as G
    model, pr = _module_case()          # P=3, n=512, m=160, 256 problems, K=6: the shape test_gpu_chain runs
    B, P = len(pr["b"]), 3
    gen = torch.Generator(device=DEV).manual_seed(5)
    bg = G.sample_erdos_renyi(B, P, 0.3, DEV, gen)
    assert str(bg.device) == DEV and len(bg) == B and bg.n_graphs == B
    graphs = bg.to_networkx()
    assert all(nx.is_connected(g) for g in graphs) and int(bg.n_bridges.sum()) > 0
    G._cache.clear()
    ref = G.BatchGraph.from_graph_list(graphs, P, DEV)
    for f in ("ev_ptr", "ev_idx", "adj_ptr", "adj_idx", "deg", "graph_id"):
        assert torch.equal(getattr(bg, f).to(torch.int64), getattr(ref, f).to(torch.int64)), f
    assert (bg.max_events, bg.max_adj) == (ref.max_events, ref.max_adj)
    b, label = pr["b"].to(DEV), pr["label"].to(DEV)
    outs = []
    for gl in (bg, graphs):
        model.zero_grad()
        torch.manual_seed(9)
        Y, _ = model(b, gl)
        _, lf = gnn_dlasso_utils.compute_loss(Y, label)
        lf.backward()
        outs.append((Y.detach().clone(), model.seq_hyp.param.grad.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.isfinite(outs[0][0]).all()
