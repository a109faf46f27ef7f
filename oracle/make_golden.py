"""Mint ``tests/golden/*.npz`` from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py

TEST INFRASTRUCTURE.  Each fixture freezes the inputs (A, b, label, graph edge lists in
adjacency order, initial noise, hyper-parameter table / hypernetwork outputs, args) and the
outputs of the reference's own classes run on them: Y, last hyp, loss_mean / loss_final,
d loss_final / d seq_hyp.param -- in fp32 (the reference's precision) and in fp64 (same code
under ``torch.set_default_dtype(torch.float64)``, initial noise drawn in fp32 then cast).
The reference publishes no golden vectors of its own, so these are the pins.
"""
from __future__ import annotations

import argparse
import os
import sys

import networkx as nx
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_harness  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def make_args(**kw):
    d = dict(m=100, n=500, P=5, alpha_max=0.1, tau_max=0.99, rho_max=0.99, eta_max=0.99,
             max_penalty_threshold=0.8, penalty_reduction_factor=0.95, snr=4, batch_size=4,
             GHN_iter_num=15, DADMM_mode="diff", GHyp_hidden=100, graph_prob=0.5)
    d.update(kw)
    return argparse.Namespace(**d)


def graph_to_adj(g, P):
    """Adjacency lists in ``graph.neighbors`` order, flattened (ptr, idx) -- enough to rebuild an
    nx.Graph with the same neighbour iteration order."""
    ptr, idx = [0], []
    for p in range(P):
        idx.extend(list(g.neighbors(p)))
        ptr.append(len(idx))
    return np.asarray(ptr, np.int32), np.asarray(idx, np.int32)


def edges_in_insertion_order(g):
    return np.asarray(list(g.edges()), np.int32).reshape(-1, 2)


def bridged_er(P, prob, seed):
    """Driver's per-sample graph recipe, gnn_dlasso_progressive.py:181-191 (seeded here)."""
    g = nx.erdos_renyi_graph(P, max(prob, 0.3), seed=seed)
    if not nx.is_connected(g):
        comps = list(nx.connected_components(g))
        for i in range(len(comps) - 1):
            g.add_edge(list(comps[i])[0], list(comps[i + 1])[0])
    return g


def run_model1(ref, A, b, label, graphs, param, args, dtype, training, seed):
    with ref_harness.default_dtype(dtype):
        model = ref.DLASSO_unfolded(A.to(dtype), args)
        model.train(training)
        with torch.no_grad():
            model.seq_hyp.param.copy_(param.to(dtype))
        torch.manual_seed(seed)
        Y, hyp = model(b.to(dtype), graphs)
        utils = ref_harness.load("gnn_dlasso_utils")
        loss_mean, loss_final = utils.compute_loss(Y, label.to(dtype))
        loss_final.backward()
        return dict(Y=Y.detach().numpy(), hyp_last=hyp.detach().numpy(),
                    loss_mean=float(loss_mean.detach()), loss_final=float(loss_final.detach()),
                    dparam=model.seq_hyp.param.grad.detach().numpy())


def case_model1(name, P, n, m, K, B, mode, graphs, param, training, alpha_max=0.1, a_scale=1.0, seed_A=0, seed_noise=7):
    ref = ref_harness.load("unfolded_DLASSO")
    utils = ref_harness.load("gnn_dlasso_utils")
    data = ref_harness.load("gnn_data")
    args = make_args(m=m, n=n, P=P, GHN_iter_num=K, DADMM_mode=mode, alpha_max=alpha_max, batch_size=B)
    torch.manual_seed(seed_A)
    A = utils.set_A(args) * a_scale
    loader = data.set_Data(A, data_len=B, args=args)
    b, label = loader.dataset.b.clone(), loader.dataset.y.clone()
    torch.manual_seed(seed_noise)
    # raw N(0,1) draws of unfolded_DLASSO.py:49-51; y0 = noise_y * 1e-2 in the run's dtype
    y0 = torch.randn((B, P, n, 1))
    U0 = torch.randn((B, P, n, 1))
    d0 = torch.randn((B, P, n, 1))
    r32 = run_model1(ref, A, b, label, graphs, param, args, torch.float32, training, seed_noise)
    r64 = run_model1(ref, A, b, label, graphs, param, args, torch.float64, training, seed_noise)
    uniq, gid = [], []
    for g in graphs:
        for i, u in enumerate(uniq):
            if u is g:
                gid.append(i)
                break
        else:
            uniq.append(g)
            gid.append(len(uniq) - 1)
    out = dict(model=np.asarray(1), P=P, n=n, m=m, K=K, B=B, mode=np.asarray(mode), training=np.asarray(training),
               max_param=np.asarray([args.alpha_max, args.tau_max, args.rho_max, args.eta_max], np.float64),
               A=A.numpy(), b=b.numpy(), label=label.numpy(), noise_y=y0.numpy(), noise_U=U0.numpy(), noise_d=d0.numpy(),
               param=param.numpy(), graph_id=np.asarray(gid, np.int32), n_graphs=len(uniq),
               Y=r32["Y"], hyp_last=r32["hyp_last"], loss_mean=r32["loss_mean"],
               loss_final=r32["loss_final"], dparam=r32["dparam"],
               Y64=r64["Y"], loss_mean64=r64["loss_mean"], loss_final64=r64["loss_final"], dparam64=r64["dparam"])
    for i, g in enumerate(uniq):
        out[f"adj_ptr_{i}"], out[f"adj_idx_{i}"] = graph_to_adj(g, P)
        out[f"edges_{i}"] = edges_in_insertion_order(g)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: Y{r32['Y'].shape} loss_final={r32['loss_final']:.8f} (fp64 {r64['loss_final']:.8f}) -> {os.path.getsize(path)/1e3:.0f} KB")


def case_model3(name, P, n, m, K, B, hidden, seed=3):
    """Model #3 (stub GCNConv, eval mode): freeze the hypernetwork's per-iteration outputs
    (captured with a forward hook on ``fc``) so the recurrence can be checked independently."""
    ref = ref_harness.load("gnn_dlasso_models_progressive")
    utils = ref_harness.load("gnn_dlasso_utils")
    data = ref_harness.load("gnn_data")
    args = make_args(m=m, n=n, P=P, GHN_iter_num=K, DADMM_mode="diff", GHyp_hidden=hidden, alpha_max=0.05, batch_size=B)
    torch.manual_seed(seed)
    A = utils.set_A(args) * 0.1
    loader = data.set_Data(A, data_len=B, args=args)
    b, label = loader.dataset.b.clone(), loader.dataset.y.clone()
    graphs = [bridged_er(P, 0.4, seed=100 + i) for i in range(B)]
    torch.manual_seed(seed + 1)
    model = ref.DLASSO_GNNHyp3_Progressive(A, args)
    model.eval()
    fc_out = []
    hook = model.fc.register_forward_hook(lambda mod, inp, out: fc_out.append(out.detach().clone()))
    torch.manual_seed(11)
    # raw N(0,1) draws of unfolded_DLASSO.py:49-51; y0 = noise_y * 1e-2 in the run's dtype
    y0 = torch.randn((B, P, n, 1))
    U0 = torch.randn((B, P, n, 1))
    d0 = torch.randn((B, P, n, 1))
    torch.manual_seed(11)
    Y, (al, ta, rh, et) = model(b, graphs, training_iterations=K)
    lm, lf = utils.compute_loss(Y, label)
    lf.backward()
    hook.remove()
    # per-iteration hyper-parameters exactly as the reference derives them (:169-196)
    hyps = []
    for o in fc_out:
        h = torch.clamp(torch.sigmoid(o), min=1e-4, max=0.9999).view(B, 4, P)
        a_ = h[:, 0] * args.alpha_max
        t_ = torch.clamp(h[:, 1] * args.tau_max, max=0.9999)
        r_ = torch.clamp(h[:, 2] * args.rho_max, max=0.9999)
        e_ = torch.clamp(h[:, 3] * args.eta_max, max=0.9999)
        hyps.append(torch.stack([a_, t_, r_, e_], dim=-1))       # [B,P,4]
    sd = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    gsd = {k: (p.grad.detach().numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32))
           for k, p in model.named_parameters()}
    out = dict(model=np.asarray(3), P=P, n=n, m=m, K=K, B=B, hidden=hidden,
               max_param=np.asarray([args.alpha_max, args.tau_max, args.rho_max, args.eta_max], np.float64),
               A=A.numpy(), b=b.numpy(), label=label.numpy(), noise_y=y0.numpy(), noise_U=U0.numpy(), noise_d=d0.numpy(),
               hyp=torch.stack(hyps).numpy(), fc_out=torch.stack(fc_out).numpy(), Y=Y.detach().numpy(),
               loss_mean=float(lm.detach()), loss_final=float(lf.detach()),
               alpha_last=al.detach().numpy(), graph_id=np.arange(B, dtype=np.int32), n_graphs=B)
    for i, g in enumerate(graphs):
        out[f"adj_ptr_{i}"], out[f"adj_idx_{i}"] = graph_to_adj(g, P)
        out[f"edges_{i}"] = edges_in_insertion_order(g)
    for k, v in sd.items():
        out["sd::" + k] = v
    for k, v in gsd.items():
        out["grad::" + k] = v
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: Y{tuple(Y.shape)} loss_final={float(lf.detach()):.8f} -> {os.path.getsize(path)/1e3:.0f} KB")


def sparse_bridged_er(P, prob, seed):
    """bench.py's config-4 graph recipe: ER at the stated probability (no 0.3 floor), components bridged as in
    gnn_dlasso_progressive.py:186-190 -- sparse graphs with leaves, bridges and long neighbour-order permutations."""
    g = nx.erdos_renyi_graph(P, prob, seed=seed)
    if not nx.is_connected(g):
        comps = list(nx.connected_components(g))
        for i in range(len(comps) - 1):
            g.add_edge(list(comps[i])[0], list(comps[i + 1])[0])
    return g


def extra_cases():
    """Fixtures used by the CPU oracle tests only (tests/helpers.py: MODEL1_EXTRA_CASES)."""
    # 6. config-4-like structure in miniature: many agents, sparse bridged per-problem graphs, train-mode table with
    #    random entries (some rows above the 0.8 penalty threshold), m = n/4
    P, n, m, K, B = 12, 40, 10, 10, 6
    graphs = [sparse_bridged_er(P, 0.12, seed=300 + i) for i in range(B)]
    torch.manual_seed(6)
    case_model1("m1_cfg4like_P12_n40", P=P, n=n, m=m, K=K, B=B, mode="diff", graphs=graphs,
                param=torch.randn(K, P, 4) * 0.3 + 0.2, training=True, a_scale=0.1)


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--extra-only" in sys.argv:
        extra_cases()
        return
    # 1. untrained (zero-init) table, default set_A conditioning (chaotic regime), one shared ER graph
    g = nx.erdos_renyi_graph(5, 0.5, seed=1)
    case_model1("m1_zero_P5_n64", P=5, n=64, m=16, K=8, B=4, mode="diff", graphs=[g] * 4,
                param=torch.zeros(8, 5, 4), training=True)
    # 2. the reference's own sizes and a TRAINED table (results/csv_folder1/model.pt), shared graph
    sd = torch.load(os.path.join(ref_harness.REFERENCE_ROOT, "results/csv_folder1/model.pt"),
                    weights_only=False, map_location="cpu")
    g = nx.Graph()
    g.add_nodes_from(range(5))
    g.add_edges_from([(0, 1), (0, 2), (0, 3), (3, 4)])       # graphs data/erods_renyi/graph_data_prob0.5_P=5.npy
    case_model1("m1_trained_P5_n500", P=5, n=500, m=100, K=25, B=2, mode="diff", graphs=[g] * 2,
                param=sd["seq_hyp.param"].clone(), training=True)
    # 3. 'same' mode, per-sample bridged graphs (non-ascending adjacency order), well-conditioned A (sigma=1)
    graphs = [bridged_er(8, 0.3, seed=40 + i) for i in range(5)]
    torch.manual_seed(5)
    case_model1("m1_same_pergraph_P8_n48", P=8, n=48, m=12, K=6, B=5, mode="same", graphs=graphs,
                param=torch.randn(6, 1, 4) * 0.3, training=False, a_scale=0.1)
    # 4. odd sizes: n not a multiple of 4, P prime, eval mode, K=15 trained table of the cuda run
    sd = torch.load(os.path.join(ref_harness.REFERENCE_ROOT,
                                 "results/P_5_num_epoch_220_train_100_test_64_batch_64_GHN_iter_num_15_lr_4e-3/model.pt"),
                    weights_only=False, map_location="cpu")
    g = nx.erdos_renyi_graph(5, 0.5, seed=9)
    case_model1("m1_trained15_P5_n51", P=5, n=51, m=13, K=15, B=3, mode="diff", graphs=[g] * 3,
                param=sd["seq_hyp.param"].clone(), training=False, alpha_max=0.06, a_scale=0.1)
    # 5. model #3 recurrence with frozen hypernetwork outputs
    case_model3("m3_frozen_P5_n32", P=5, n=32, m=8, K=4, B=3, hidden=8)
    extra_cases()


if __name__ == "__main__":
    main()
