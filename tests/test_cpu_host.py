"""CPU-side tests (-m "not gpu"): the C-ABI library loads and exports every declared symbol, host logic
(graph CSR, hyper-parameter table, drop-in module surface, sharding) -- no compute calls."""
import argparse
import os
import re

import networkx as nx
import numpy as np
import pytest
import torch

from helpers import MODEL1_CASES, ROOT, PKG, Golden
from oracle import dadmm_oracle as O


def test_library_loads_and_exports_header_symbols():
    import ctypes
    from dadmm_b200 import _lib
    header = open(os.path.join(ROOT, "include", "dadmm.h")).read()
    declared = set(re.findall(r"\b(dadmm_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 17
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/dadmm.h but not exported"
    assert set(_lib.EXPORTED) == declared
    assert _lib.lib.dadmm_abi_version() == _lib.ABI_VERSION == 7


def test_invalid_arguments_return_error_codes_without_a_gpu():
    from dadmm_b200 import _lib
    rc = _lib.lib.dadmm_contract(0, 0, 0, 1, 1, 1, None, 0, 0, 0, None, 0, 0, 0, None, 0, 0, 0, 0, None, 0, None)
    assert rc < 0 and b"bad dims" in _lib.lib.dadmm_last_error()
    rc = _lib.lib.dadmm_contract(0, 0, 2, 1, 4, 4, None, 0, 0, 0, None, 0, 0, 0, None, 0, 0, 0, 0, None, 0, None)
    assert rc < 0 and b"null pointer" in _lib.lib.dadmm_last_error()
    assert _lib.lib.dadmm_partials_elems(0, 4, 5, 100) == 4 * 4 * 5 * 4
    assert _lib.lib.dadmm_unfolded_ws_bytes(0, 1, 2, 3, 8, 4, 0, 0) >= 3 * 2 * 3 * 8 * 4
    assert _lib.lib.dadmm_unfolded_ws_bytes(0, 1, 3, 5, 51, 15, 1, 0) % 256 == 0


def test_cpu_tensors_are_rejected_loudly():
    from dadmm_b200 import functional as DF
    from dadmm_b200._lib import DadmmError
    with pytest.raises(DadmmError, match="no CPU fallback"):
        DF.contract(torch.zeros(1, 4, 4), torch.zeros(2, 1, 4))


@pytest.mark.parametrize("name", MODEL1_CASES + ["m3_frozen_P5_n32"])
def test_event_csr_matches_oracle(name):
    from dadmm_b200.graph import BatchGraph
    g = Golden(name)
    h = BatchGraph.build_host(g.graphs, g.P)
    ptr, idx, deg, gid, G = h.ev_ptr, h.ev_idx, h.deg, h.graph_id, h.n_graphs
    uniq = []
    for gr in g.graphs:
        if not any(gr is u for u in uniq):
            uniq.append(gr)
    assert G == len(uniq) and len(ptr) == G * g.P + 1
    for gi, gr in enumerate(uniq):
        ev = O.event_lists(gr, g.P)
        for p in range(g.P):
            node = gi * g.P + p
            assert idx[ptr[node]:ptr[node + 1]].tolist() == ev[p]
            assert deg[node] == len(list(gr.neighbors(p)))
            assert len(ev[p]) == 2 * deg[node]
            assert h.adj_idx[h.adj_ptr[node]:h.adj_ptr[node + 1]].tolist() == list(gr.neighbors(p))
    if gid is not None:
        assert [uniq[i] is gr for i, gr in zip(gid, g.graphs)] == [True] * g.B


def test_self_loop_and_isolated_node():
    from dadmm_b200.graph import BatchGraph
    gr = nx.Graph()
    gr.add_nodes_from(range(4))
    gr.add_edges_from([(0, 1), (1, 1), (1, 2)])          # node 3 isolated, self-loop on 1
    h = BatchGraph.build_host([gr, gr], 4)
    ptr, idx, deg, gid, G = h.ev_ptr, h.ev_idx, h.deg, h.graph_id, h.n_graphs
    assert gid is None and G == 1
    assert h.adj_idx[h.adj_ptr[1]:h.adj_ptr[2]].tolist() == [0, 2] and h.max_adj == 4 and h.max_events == 10
    assert deg.tolist() == [1, 3, 1, 0]
    assert idx[ptr[3]:ptr[4]].tolist() == []
    y = torch.randn(1, 4, 6, 1)
    d = O.delta_events([gr], y)
    assert torch.allclose(d[0, 1], 2 * ((y[0, 1] - y[0, 0]) + (y[0, 1] - y[0, 2])), atol=1e-6)
    assert torch.equal(d[0, 3], torch.zeros(6, 1))


def _args(P=5, n=8, m=4, K=6, mode="diff", **kw):
    d = dict(m=m, n=n, P=P, GHN_iter_num=K, DADMM_mode=mode, alpha_max=0.1, tau_max=0.99, rho_max=0.99, eta_max=0.99,
             max_penalty_threshold=0.8, penalty_reduction_factor=0.95, batch_size=2, snr=4, GHyp_hidden=4)
    d.update(kw)
    return argparse.Namespace(**d)


@pytest.mark.parametrize("mode", ["diff", "same"])
@pytest.mark.parametrize("training", [True, False])
def test_hyp_table_matches_oracle(mode, training):
    """forward(k) is bit-identical to the reference recipe; the all-K table differs from it only by the
    CPU vectoriser's sigmoid rounding (<= 1 ulp) -- on CUDA both are the same element-wise kernel."""
    import unfolded_DLASSO
    args = _args(mode=mode)
    model = unfolded_DLASSO.DLASSO_unfolded(torch.randn(1, 5, 4, 8), args)
    model.train(training)
    with torch.no_grad():
        model.seq_hyp.param.copy_(torch.randn_like(model.seq_hyp.param) * 1.5 + 1.0)   # some rows above the 0.8 threshold
    table = model.seq_hyp.table(6)
    ref = O.hyp_table(model.seq_hyp.param.detach(), torch.tensor([0.1, 0.99, 0.99, 0.99]), training)
    assert torch.allclose(table.detach(), ref, rtol=0, atol=1.2e-7)
    for k in (0, 3, 5):
        assert torch.equal(model.seq_hyp(k).detach(), ref[k].unsqueeze(-1))
    assert list(model.state_dict().keys()) == ["seq_hyp.param"]
    for attr in ("A", "P", "m", "n", "K", "DADMM_mode", "seq_hyp", "max_param", "args"):
        assert hasattr(model, attr)


@pytest.mark.ref
def test_seq_hyperparam_matches_reference_class():
    from oracle import ref_harness
    import unfolded_DLASSO
    ref = ref_harness.load("unfolded_DLASSO")
    args = _args()
    mp = torch.tensor([0.1, 0.99, 0.99, 0.99])
    ours, theirs = unfolded_DLASSO.seq_hyperparam([6, 5, 4], mp, args), ref.seq_hyperparam([6, 5, 4], mp, args)
    p = torch.randn(6, 5, 4) * 1.5 + 1.0
    with torch.no_grad():
        ours.param.copy_(p)
        theirs.param.copy_(p)
    for training in (True, False):
        ours.train(training)
        theirs.train(training)
        for k in range(6):
            assert torch.equal(ours(k), theirs(k))


@pytest.mark.ref
def test_reference_checkpoints_load():
    from oracle import ref_harness
    import unfolded_DLASSO
    root = ref_harness.REFERENCE_ROOT
    for rel in ("results/csv_folder1", "results/P_5_num_epoch_220_train_100_test_64_batch_64_GHN_iter_num_15_lr_4e-3"):
        args = torch.load(os.path.join(root, rel, "args.pt"), weights_only=False)
        A = torch.load(os.path.join(root, rel, "A.pt"), weights_only=False, map_location="cpu")
        model = unfolded_DLASSO.DLASSO_unfolded(A, args)
        model.load_state_dict(torch.load(os.path.join(root, rel, "model.pt"), weights_only=False, map_location="cpu"))
        assert model.seq_hyp.param.shape == (args.GHN_iter_num, args.P, 4)


@pytest.mark.ref
def test_configurations_match_reference_defaults(monkeypatch):
    from oracle import ref_harness
    import configurations
    monkeypatch.setattr("sys.argv", ["prog"])
    ref = ref_harness.load("configurations").args_parser()
    ours = configurations.args_parser([])
    assert vars(ours) == vars(ref)


@pytest.mark.ref
def test_set_A_and_set_Data_match_reference_rng_stream():
    from oracle import ref_harness
    import gnn_dlasso_utils
    import gnn_data
    args = _args(P=3, n=12, m=5)
    torch.manual_seed(0)
    A_ref = ref_harness.load("gnn_dlasso_utils").set_A(args)
    ld_ref = ref_harness.load("gnn_data").set_Data(A_ref, 6, args)
    tail_ref = torch.rand(1)
    torch.manual_seed(0)
    A = gnn_dlasso_utils.set_A(args)
    ld = gnn_data.set_Data(A, 6, args)
    tail = torch.rand(1)
    assert torch.equal(A, A_ref) and torch.equal(tail, tail_ref)
    assert torch.equal(ld.dataset.b, ld_ref.dataset.b) and torch.equal(ld.dataset.y, ld_ref.dataset.y)


def test_model3_module_surface_and_state_dict_keys():
    import gnn_dlasso_models_progressive as M
    g = Golden("m3_frozen_P5_n32")
    args = _args(P=g.P, n=g.n, m=g.m, K=g.K, GHyp_hidden=float(g.z["hidden"]))     # the flag arrives as float upstream
    model = M.DLASSO_GNNHyp3_Progressive(g.t("A"), args)
    sd = {k[4:]: torch.from_numpy(g.z[k]) for k in g.z.files if k.startswith("sd::")}
    assert set(model.state_dict().keys()) == set(sd.keys()) and len(sd) == 51
    model.load_state_dict(sd)
    assert model.fc.bias.shape == (4 * g.P,)


@pytest.mark.ref
def test_model3_init_matches_reference_under_same_seed():
    """Same RNG consumption at construction as the reference (+ PyG-semantics stub): identical initial weights."""
    from oracle import ref_harness
    import gnn_dlasso_models_progressive as M
    ref = ref_harness.load("gnn_dlasso_models_progressive")
    args = _args(P=4, n=6, m=3, K=2, GHyp_hidden=4)
    A = torch.randn(1, 4, 3, 6)
    torch.manual_seed(5)
    theirs = ref.DLASSO_GNNHyp3_Progressive(A, args)
    torch.manual_seed(5)
    ours = M.DLASSO_GNNHyp3_Progressive(A, args)
    sd_o, sd_t = ours.state_dict(), theirs.state_dict()
    assert sd_o.keys() == sd_t.keys()
    for k in sd_o:
        assert torch.equal(sd_o[k], sd_t[k]), k


def test_batched_gcn_encoder_matches_per_sample_stub():
    """Hypernetwork encoder: batched dense GCN (ours) vs per-sample PyG-semantics stub, eval AND train mode
    (train mode without dropout: checks per-sample BatchNorm statistics and the closed-form running-stat update)."""
    from oracle import ref_harness
    import gnn_dlasso_models_progressive as M
    if not ref_harness.reference_available():
        pytest.skip("reference checkout absent")
    ref = ref_harness.load("gnn_dlasso_models_progressive")
    P, m, h, B = 6, 10, 4, 5
    torch.manual_seed(2)
    theirs = ref.GNNHypernetwork3(P, m, h)
    ours = M.GNNHypernetwork3(P, m, h)
    ours.load_state_dict(theirs.state_dict())
    graphs = [nx.erdos_renyi_graph(P, 0.5, seed=i) for i in range(B)]
    x = torch.randn(B, P, m, 1)
    for mod in (ours, theirs):
        mod.dropout.p = 0.0
    for training in (True, False):
        ours.train(training)
        theirs.train(training)
        a, b = ours(x, graphs), theirs(x, graphs)
        assert torch.allclose(a, b, atol=2e-5, rtol=1e-4), float((a - b).abs().max())
        for k, v in ours.state_dict().items():
            assert torch.allclose(v.float(), theirs.state_dict()[k].float(), atol=1e-5, rtol=1e-5), k
        # the reference's per-sample entry point (:42-72) exists here too and takes the same path with a batch of one
        a1, b1 = ours.graph_conv(x[0], graphs[0]), theirs.graph_conv(x[0], graphs[0])
        assert a1.shape == b1.shape == (P * 4 * h,) and torch.allclose(a1, b1, atol=5e-5, rtol=1e-4), float((a1 - b1).abs().max())
        for k, v in ours.state_dict().items():
            assert torch.allclose(v.float(), theirs.state_dict()[k].float(), atol=1e-5, rtol=1e-5), k


def test_shard_ranges_cover_batch():
    from dadmm_b200 import dist as D
    for B in (1, 7, 8, 4096):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


# ------------------------------------------------------------------------------------------ graph ingestion
def _ingestion_graphs(P):
    import networkx as nx
    gs = [nx.erdos_renyi_graph(P, 0.3, seed=s) for s in range(12)]
    odd = nx.Graph()
    odd.add_nodes_from(range(P))
    odd.add_edges_from([(0, 1), (1, 1), (3, 2), (2, 0), (P - 1, 0)])      # a self-loop, isolated nodes, unsorted insertion order
    empty = nx.Graph()
    empty.add_nodes_from(range(P))
    return gs + [odd, empty, gs[0], odd]                                   # repeated objects are de-duplicated


@pytest.mark.parametrize("builder", ["numpy", "torch"])
def test_vectorised_graph_ingestion_equals_per_graph_event_lists(builder):
    """HostGraph (numpy passes) and _DeviceCSR (torch passes; on the CPU device here, on cuda in the gpu tests) against
    the per-graph construction that mirrors unfolded_DLASSO.py:111-118,132-139 line by line."""
    from dadmm_b200 import graph as G
    P = 9
    gs = _ingestion_graphs(P)
    h = G.HostGraph(gs, P) if builder == "numpy" else G._DeviceCSR(gs, P, "cpu")
    get = lambda a: np.asarray(a.cpu() if hasattr(a, "cpu") else a)
    ev_ptr, ev_idx, adj_ptr, adj_idx, deg = (get(x) for x in (h.ev_ptr, h.ev_idx, h.adj_ptr, h.adj_idx, h.deg))
    assert h.n_graphs == 14 and list(h.graph_id[-2:]) == [0, 12]
    worst_ev = worst_adj = 0
    for gi, g in enumerate(h.unique_graphs):
        ev = G.event_lists(g, P)
        worst_ev = max(worst_ev, sum(map(len, ev)))
        nadj = 0
        for p in range(P):
            q = gi * P + p
            assert ev_idx[ev_ptr[q]:ev_ptr[q + 1]].tolist() == ev[p]
            nb = list(g.neighbors(p))
            assert deg[q] == len(nb)
            assert adj_idx[adj_ptr[q]:adj_ptr[q + 1]].tolist() == [j for j in nb if j != p]
            nadj += len([j for j in nb if j != p])
        worst_adj = max(worst_adj, nadj)
    assert (h.max_events, h.max_adj) == (worst_ev, worst_adj)
    assert ev_idx.dtype == np.int32 and ev_ptr.dtype == np.int32


def test_graph_ingestion_rejects_foreign_nodes():
    import networkx as nx
    from dadmm_b200 import graph as G
    g = nx.Graph()
    g.add_nodes_from(range(4))
    g.add_edge(0, 7)
    with pytest.raises(ValueError, match="outside 0..3"):
        G.HostGraph([g], 4)
    h = nx.Graph()
    h.add_nodes_from(range(3))
    with pytest.raises(ValueError, match="no node 3"):
        G.HostGraph([h], 4)


def test_hyper_parameter_table_backward_closed_form_matches_autograd():
    """seq_hyperparam.table: rows bit-identical to the per-k torch.sum of the reference (unfolded_DLASSO.py:157); its
    backward (one masked product + one reduction) equals autograd through the K slices, also when K < len(param)."""
    import argparse
    import unfolded_DLASSO as U
    args = argparse.Namespace(max_penalty_threshold=0.8, penalty_reduction_factor=0.95)
    mp = torch.tensor([0.1, 0.99, 0.99, 0.99])
    for shape, K in (([25, 50, 4], 25), ([15, 1, 4], 15), ([15, 5, 4], 9)):
        torch.manual_seed(K)
        sh = U.seq_hyperparam(shape, mp, args).double()
        with torch.no_grad():
            sh.param.copy_(torch.randn(shape) * 0.7)
        w = torch.randn((K, shape[1], 4), dtype=torch.float64)
        t = sh.table(K)
        (t * w).sum().backward()
        g_closed, sh.param.grad = sh.param.grad.clone(), None
        t_ref = sh._squash(torch.stack([torch.sum(sh.param[:k + 1], dim=0) for k in range(K)]))
        (t_ref * w).sum().backward()
        assert torch.equal(t, t_ref)
        assert g_closed.shape == sh.param.shape
        assert float((g_closed - sh.param.grad).abs().max()) < 1e-14
        if K < shape[0]:
            assert float(g_closed[K:].abs().max()) == 0.0


def test_per_sample_batchnorm_function_matches_composed_formula_and_gradcheck():
    """_PerSampleBatchNorm (one autograd node, textbook backward) against the composed element-wise formula it replaces
    and against numerical differentiation, fp64."""
    import gnn_dlasso_models_progressive as M
    torch.manual_seed(3)
    B, P, Cn, eps = 4, 6, 5, 1e-5
    x = torch.randn(B, P, Cn, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Cn, dtype=torch.float64, requires_grad=True)
    b = torch.randn(Cn, dtype=torch.float64, requires_grad=True)
    out, mean, var = M._PerSampleBatchNorm.apply(x, w, b, eps)
    mean_r = x.mean(dim=1, keepdim=True)
    var_r = x.var(dim=1, unbiased=False, keepdim=True)
    ref = (x - mean_r) / torch.sqrt(var_r + eps) * w + b
    assert torch.allclose(out, ref, atol=1e-12) and torch.allclose(mean, mean_r, atol=1e-14) and torch.allclose(var, var_r, atol=1e-14)
    assert not mean.requires_grad and not var.requires_grad
    g = torch.randn_like(out)
    got = torch.autograd.grad(out, (x, w, b), g, retain_graph=True)
    want = torch.autograd.grad(ref, (x, w, b), g)
    for a, r in zip(got, want):
        assert torch.allclose(a, r, atol=1e-10), float((a - r).abs().max())
    assert torch.autograd.gradcheck(lambda *a: M._PerSampleBatchNorm.apply(*a, eps)[0], (x, w, b), eps=1e-6, atol=1e-6)


@pytest.mark.parametrize("mode", ["diff", "same"])
def test_model3_packed_hyperparameter_scaling_equals_reference_formula(mode):
    """One multiply + one clamp over the packed [B,4,P|1,1,1] tensor == the reference's four multiplies and three clamps
    (gnn_dlasso_models_progressive.py:170-196), values bit for bit and gradients."""
    import gnn_dlasso_models_progressive as M
    P, n, m, B = 4, 6, 3, 7
    args = _args(P=P, n=n, m=m, K=2, GHyp_hidden=4)
    args.DADMM_mode = mode
    torch.manual_seed(11)
    model = M.DLASSO_GNNHyp3_Progressive(torch.randn(1, P, m, n), args)
    raw = (torch.randn(B, 4 * (P if mode == "diff" else 1)) * 6).requires_grad_(True)      # sigmoid saturates on some entries
    packed = model._scaled(raw, B)
    assert packed.shape == (B, 4, P if mode == "diff" else 1, 1, 1)
    h = torch.clamp(torch.sigmoid(raw), min=1e-4, max=0.9999).view(B, 4, -1, 1, 1)
    am, tm, rm, em = (float(t) for t in (model.alpha_max, model.tau_max, model.rho_max, model.eta_max))
    ref = (h[:, 0] * am, torch.clamp(h[:, 1] * tm, max=0.9999), torch.clamp(h[:, 2] * rm, max=0.9999),
           torch.clamp(h[:, 3] * em, max=0.9999))
    for a, r in zip(packed.unbind(dim=1), ref):
        assert torch.equal(a, r)
    wts = torch.randn(B, 4, packed.shape[2], 1, 1)
    g1, = torch.autograd.grad((packed * wts).sum(), raw, retain_graph=True)
    g2, = torch.autograd.grad(sum((r * wts[:, i]).sum() for i, r in enumerate(ref)), raw)
    assert torch.allclose(g1, g2, atol=1e-7)


def test_vectorised_normalized_adjacency_equals_per_graph_construction():
    """GCN propagation matrices D^-1/2 (Adj + I) D^-1/2 (gnn_dlasso_models_progressive.py:43 via PyG's gcn_norm): the
    all-graphs-at-once construction (module function and the BatchGraph-cached form the forward uses) against the
    per-graph loop it replaces, bit for bit -- self-loops, isolated nodes, repeated graph objects, one shared graph."""
    import networkx as nx
    import gnn_dlasso_models_progressive as M
    from dadmm_b200.graph import BatchGraph

    def per_graph(graph_list, P):
        out = []
        for g in graph_list:
            a = np.eye(P)
            for u, v in g.edges():
                if u != v:
                    a[u, v] = a[v, u] = 1.0
            d = a.sum(axis=0) ** -0.5
            out.append(d[:, None] * a * d[None, :])
        return torch.from_numpy(np.stack(out)).float()

    P = 9
    gs = _ingestion_graphs(P)
    want = per_graph(gs, P)
    assert torch.equal(M.normalized_adjacency(gs, P, "cpu"), want)
    bg = BatchGraph.from_graph_list(gs, P, "cpu")
    got = bg.normalized_adjacency(torch.float32)
    assert torch.equal(got, want) and bg.normalized_adjacency(torch.float32) is got          # cached
    assert torch.equal(M.normalized_adjacency(bg, P, "cpu"), want)
    assert torch.equal(bg.shard(3, 9).normalized_adjacency(), want[3:9])
    one = [gs[2]] * 5
    assert torch.equal(BatchGraph.from_graph_list(one, P, "cpu").normalized_adjacency(), per_graph(one, P))
    assert torch.equal(M.normalized_adjacency(one, P, "cpu"), per_graph(one, P))


def test_abi_v5_host_side_contracts_without_a_gpu():
    """Pieces of ABI v5 that are decided on the host: the PDL switch, operator-split sizes per route, the reverse
    sweep's partial-sum workspace, and the rule for omitting Atb."""
    import ctypes as C
    from dadmm_b200 import _lib
    lib = _lib.lib
    prev = _lib.set_pdl(False)
    assert _lib.set_pdl(True) is False and _lib.set_pdl(prev) is True
    # operator split: nothing off the fused tensor-core path; W alone on the single-stage route; F1 + F2 on the two-stage one
    B, P, n, m = 256, 3, 512, 160
    split = lambda rows, k: 256 + (rows * ((k + 7) // 8 * 8) * 4 + 255) // 256 * 256          # f16::split_bytes
    assert lib.dadmm_unfolded_op_split_bytes(_lib.F64, _lib.ALGO_AUTO, B, P, n, m) == 0
    assert lib.dadmm_unfolded_op_split_bytes(_lib.F32, _lib.ALGO_SIMT, B, P, n, m) == 0
    assert lib.dadmm_unfolded_op_split_bytes(_lib.F32, _lib.ALGO_AUTO, 8, P, n, m) == 0            # batch below a tile
    assert lib.dadmm_unfolded_uses_factor(_lib.F32, _lib.ALGO_AUTO, B, P, n, m) == 1
    assert lib.dadmm_unfolded_op_split_bytes(_lib.F32, _lib.ALGO_AUTO, B, P, n, m) == split(P * m, n) + split(P * n, m)
    assert lib.dadmm_unfolded_op_split_bytes(_lib.F32, _lib.ALGO_AUTO, B, P, n, 0) == split(P * n, n)
    assert lib.dadmm_unfolded_op_split_bytes(_lib.F32, _lib.ALGO_AUTO, B, P, n, 256) == split(P * n, n)   # 8m > 3n: single stage
    # reverse sweep: K levels of [csplit][B][P][4] partial rows fit the workspace
    K = 6
    fwd, bwd = (lib.dadmm_unfolded_ws_bytes(_lib.F32, _lib.ALGO_SIMT, B, P, n, K, back, 0) for back in (0, 1))
    assert bwd >= 3 * B * P * n * 4 + K * B * P * 4 * 4 and fwd % 256 == 0 and bwd % 256 == 0
    # Atb may be omitted only on the two-stage route with the observation term in the factor
    clamps = (_lib.Clamps * K)()
    one = C.c_void_p(256)                                  # never dereferenced: validation fails first
    g = _lib.Graph(1, P, 256, 256, 256, None, 256, 256, 0, 0)

    def fwd_rc(atb, fac):
        return lib.dadmm_unfolded_fwd(_lib.F32, _lib.ALGO_AUTO, B, P, n, K, C.byref(g), clamps, one, one,
                                      C.byref(fac) if fac else None, atb, one, one, one, one, None, None, one, 0, None, None, None, None)
    assert fwd_rc(None, None) < 0 and b"Atb is required" in lib.dadmm_last_error()
    assert fwd_rc(None, _lib.Factor(m, 256, 256, None)) < 0 and b"Atb is required" in lib.dadmm_last_error()
    assert fwd_rc(None, _lib.Factor(256, 256, 256, 256)) < 0 and b"Atb is required" in lib.dadmm_last_error()   # not two-stage
    rc = fwd_rc(None, _lib.Factor(m, 256, 256, 256))      # allowed: the next check (workspace size 0) is the one that fires
    assert rc < 0 and b"workspace too small" in lib.dadmm_last_error()


def test_graph_ingestion_property_random_insertion_orders():
    """Property test (hypothesis): for random multigraph-free edge lists in random insertion order -- self-loops, isolated
    nodes and duplicates included -- the vectorised CSR equals the per-graph event lists of unfolded_DLASSO.py:132-139,
    degrees equal ``len(list(g.neighbors(p)))``, and delta = 2 L y computed from the event lists equals the reference's
    double loop."""
    import networkx as nx
    from hypothesis import given, settings, strategies as st
    from dadmm_b200 import graph as G

    @settings(max_examples=40, deadline=None)
    @given(st.integers(1, 9).flatmap(lambda P: st.tuples(st.just(P), st.lists(st.lists(
        st.tuples(st.integers(0, P - 1), st.integers(0, P - 1)), max_size=14), min_size=1, max_size=4))))
    def check(case):
        P, edge_lists = case
        gs = []
        for edges in edge_lists:
            g = nx.Graph()
            g.add_nodes_from(range(P))
            g.add_edges_from(edges)
            gs.append(g)
        h = G.HostGraph(gs, P)
        rng = np.random.default_rng(P)
        y = rng.standard_normal((len(gs), P, 3))
        for gi, g in enumerate(gs):
            ev = G.event_lists(g, P)
            delta = np.zeros((P, 3))
            for p in range(P):                       # the reference's accumulation, unfolded_DLASSO.py:132-139
                for j in g.neighbors(p):
                    diff = y[gi, p] - y[gi, j]
                    delta[p] += diff
                    delta[j] -= diff
            for p in range(P):
                q = gi * P + p
                lst = h.ev_idx[h.ev_ptr[q]:h.ev_ptr[q + 1]].tolist()
                assert lst == ev[p]
                assert h.deg[q] == len(list(g.neighbors(p)))
                acc = np.zeros(3)
                for e in lst:                        # what the forward level kernel does with the event list
                    acc = acc + (y[gi, p] - y[gi, e])
                assert np.array_equal(acc, delta[p])
    check()


# ---------------------------------------------------------------------------------------------------------
# batched Erdos-Renyi sampler: same tensor passes on the CPU here as on the GPU
# ---------------------------------------------------------------------------------------------------------
_CSR_FIELDS = ("ev_ptr", "ev_idx", "adj_ptr", "adj_idx", "deg", "graph_id")


@pytest.mark.parametrize("B,P,p", [(48, 12, 0.12), (64, 5, 0.5), (16, 30, 0.05), (8, 7, 0.0), (8, 6, 1.0), (5, 1, 0.5), (1, 9, 0.3)])
def test_sampled_batch_is_what_ingesting_its_networkx_graphs_gives(B, P, p):
    """sample_erdos_renyi builds the CSR arrays without ever making a networkx object; rebuilding the networkx graphs
    from its neighbour lists (``to_networkx``: G(P,p) edges in lexicographic order, then the bridges) and ingesting
    them the usual way gives the same arrays, every graph is connected, and the bridges are exactly the reference
    driver's chain over consecutive components (gnn_dlasso_progressive.py:186-190) with the smallest node of each."""
    from dadmm_b200 import graph as G
    bg = G.sample_erdos_renyi(B, P, p, "cpu", torch.Generator().manual_seed(1))
    raw = G.sample_erdos_renyi(B, P, p, "cpu", torch.Generator().manual_seed(1), connect=False)   # same draw, no bridges
    graphs, raw_graphs = bg.to_networkx(), raw.to_networkx()
    assert len(bg) == B and bg.n_graphs == B and bg.P == P
    G._cache.clear()
    ref = G.BatchGraph.from_graph_list(graphs, P, "cpu")
    for f in _CSR_FIELDS:
        a, r = getattr(bg, f), getattr(ref, f)
        assert (a is None and r is None) or torch.equal(a.to(torch.int64), r.to(torch.int64)), f
    assert (bg.max_events, bg.max_adj) == (ref.max_events, ref.max_adj)
    for b in range(B):
        g, r = graphs[b], raw_graphs[b]
        assert nx.is_connected(g) and g.number_of_nodes() == P
        comps = sorted(nx.connected_components(r), key=min)
        want = {(min(comps[i]), min(comps[i + 1])) for i in range(len(comps) - 1)}
        assert {tuple(sorted(e)) for e in g.edges()} - {tuple(sorted(e)) for e in r.edges()} == want
        assert int(bg.n_bridges[b]) == len(comps) - 1 and int(raw.n_bridges[b]) == 0
        # the order networkx itself would hold after the driver's construction
        d = nx.Graph()
        d.add_nodes_from(range(P))
        d.add_edges_from(sorted(tuple(sorted(e)) for e in r.edges()))
        d.add_edges_from(sorted(want))
        assert [list(d.neighbors(u)) for u in range(P)] == [list(g.neighbors(u)) for u in range(P)]
    # propagation matrices of model #3 come out of the same lists
    import gnn_dlasso_models_progressive as M
    assert torch.equal(bg.normalized_adjacency(), M.normalized_adjacency(graphs, P, "cpu"))


def test_sampler_statistics_determinism_and_argument_checks():
    from dadmm_b200 import graph as G
    B, P, p = 4000, 10, 0.3
    a = G.sample_erdos_renyi(B, P, p, "cpu", torch.Generator().manual_seed(7), connect=False)
    b = G.sample_erdos_renyi(B, P, p, "cpu", torch.Generator().manual_seed(7), connect=False)
    c = G.sample_erdos_renyi(B, P, p, "cpu", torch.Generator().manual_seed(8), connect=False)
    assert all(torch.equal(getattr(a, f), getattr(b, f)) for f in _CSR_FIELDS)
    assert not torch.equal(a.deg, c.deg)
    pairs = B * P * (P - 1) // 2
    edges = int(a.deg.sum()) // 2
    assert abs(edges - p * pairs) < 5 * (pairs * p * (1 - p)) ** 0.5              # binomial, 5 sigma
    per_node = a.deg.view(B, P).double().mean(dim=0) / (P - 1)                       # no node position is favoured
    assert float((per_node - p).abs().max()) < 5 * (p * (1 - p) / (B * (P - 1))) ** 0.5
    # isolated nodes get chained 0-1-2-...; the complete graph needs nothing
    chain = G.sample_erdos_renyi(3, 6, 0.0, "cpu")
    assert [sorted(map(sorted, g.edges())) for g in chain.to_networkx()] == [[[i, i + 1] for i in range(5)]] * 3
    full = G.sample_erdos_renyi(2, 6, 1.0, "cpu")
    assert int(full.n_bridges.sum()) == 0 and bool((full.deg == 5).all())
    for bad in ((0, 5, 0.5), (4, 0, 0.5), (4, 5, 1.5), (4, 5, -0.1)):
        with pytest.raises(ValueError):
            G.sample_erdos_renyi(*bad, "cpu")


def test_to_networkx_round_trips_arbitrary_insertion_orders():
    """``to_networkx`` recovers an insertion order for ANY graph this module ingested (not only sampled ones): shuffled
    edge insertion, self-loops, isolated nodes, shared graph objects."""
    import random
    from dadmm_b200 import graph as G
    rnd = random.Random(3)
    P, graphs = 9, []
    for s in range(12):
        g0 = nx.erdos_renyi_graph(P, 0.35, seed=s)
        e = list(g0.edges()) + ([(2, 2)] if s % 3 == 0 else [])
        rnd.shuffle(e)
        e = [(v, u) if rnd.random() < 0.5 else (u, v) for u, v in e]
        g = nx.Graph()
        g.add_nodes_from(range(P))
        g.add_edges_from(e)
        graphs.append(g)
    graphs += [graphs[0], graphs[3]]
    G._cache.clear()
    bg = G.BatchGraph.from_graph_list(graphs, P, "cpu")
    back = bg.to_networkx()
    assert back[0] is back[-2] and back[3] is back[-1]
    for g, r in zip(graphs, back):
        assert [list(g.neighbors(u)) for u in range(P)] == [list(r.neighbors(u)) for u in range(P)]
    assert [list(x.neighbors(1)) for x in bg.to_networkx([5, 0])] == [list(graphs[5].neighbors(1)), list(graphs[0].neighbors(1))]


def test_shard_of_a_sampled_batch_keeps_its_problems():
    """Multi-GPU sharding of a sampled batch (``BatchGraph.shard``): the slice keeps its problems' graphs, bridge counts
    and propagation matrices, and shares the CSR arrays."""
    from dadmm_b200 import graph as G
    bg = G.sample_erdos_renyi(10, 4, 0.2, "cpu", torch.Generator().manual_seed(5))
    sh = bg.shard(3, 7)
    assert len(sh) == 4 and sh.ev_idx.data_ptr() == bg.ev_idx.data_ptr()
    assert torch.equal(sh.n_bridges, bg.n_bridges[3:7]) and torch.equal(sh.graph_id, bg.graph_id[3:7])
    assert torch.equal(sh.normalized_adjacency(), bg.normalized_adjacency()[3:7])
    assert [sorted(g.edges()) for g in sh.to_networkx()] == [sorted(g.edges()) for g in bg.to_networkx()[3:7]]
