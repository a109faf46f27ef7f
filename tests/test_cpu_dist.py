"""world_size-2 gloo test of the N>1 host logic: batch sharding + gradient / loss all-reduce reproduce the
single-process result (per-rank compute is emulated with the CPU oracle -- the CUDA kernels are covered by
the -m gpu tests; what is checked here is the sharding arithmetic and the collectives)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT, PKG, random_problem


def _worker(rank, world, port, out):
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    from dadmm_b200 import dist as D
    from oracle import dadmm_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    P, n, m, B, K = 4, 12, 5, 6, 3
    pr = random_problem(P, n, m, B, K, seed=3)
    b, label, graphs, (lo, hi) = D.shard_batch(pr["b"], pr["label"], pr["graphs"], rank, world)
    param = pr["param"].clone().requires_grad_(True)
    table = O.hyp_table(param, torch.tensor([0.1, 0.99, 0.99, 0.99]), True)
    AtA, Atb = O.atx(pr["A"], pr["A"]), O.atx(pr["A"], b)
    Y = O.unfolded_forward(AtA, Atb, graphs, pr["y0"][lo:hi], pr["U0"][lo:hi], pr["d0"][lo:hi], table)
    # local share of the global mean loss: sum over local problems / (P * B_global * n)
    losses = ((Y - label.unsqueeze(0).unsqueeze(2)) ** 2).sum(dim=(1, 2, 3, 4)) / (P * B * n)
    losses[-1].backward()
    D.allreduce_sum_([param.grad])
    tot = losses.detach().clone()
    D.allreduce_sum_([tot])
    if rank == 0:
        torch.save(dict(grad=param.grad, losses=tot), out)
    dist.destroy_process_group()


def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_sharding_matches_single_process(tmp_path):
    from oracle import dadmm_oracle as O
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    P, n, m, B, K = 4, 12, 5, 6, 3
    pr = random_problem(P, n, m, B, K, seed=3)
    param = pr["param"].clone().requires_grad_(True)
    table = O.hyp_table(param, torch.tensor([0.1, 0.99, 0.99, 0.99]), True)
    Y = O.unfolded_forward(O.atx(pr["A"], pr["A"]), O.atx(pr["A"], pr["b"]), pr["graphs"], pr["y0"], pr["U0"], pr["d0"], table)
    losses = ((Y - pr["label"].unsqueeze(0).unsqueeze(2)) ** 2).sum(dim=(1, 2, 3, 4)) / (P * B * n)
    losses[-1].backward()
    assert torch.allclose(got["losses"], losses.detach(), rtol=1e-5, atol=1e-7)
    assert torch.allclose(got["grad"], param.grad, rtol=1e-4, atol=1e-7)


def _worker_uneven(rank, world, port, out):
    """Uneven shards (7 problems over 3 ranks), a module's gradients and the loss shares in ONE bucket
    (``allreduce_gradients(module, extra=...)``), mixed dtypes in the bucket, a parameter without a gradient."""
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    from dadmm_b200 import dist as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    B, F = 7, 5
    g = torch.Generator().manual_seed(11)
    x, t = torch.randn(B, F, generator=g), torch.randn(B, 2, generator=g)
    torch.manual_seed(5)
    net = torch.nn.Sequential(torch.nn.Linear(F, 4), torch.nn.Tanh(), torch.nn.Linear(4, 2))
    unused = torch.nn.Parameter(torch.zeros(3))                 # never reaches the loss: .grad stays None
    net.register_parameter("unused", unused)
    lo, hi = D.shard_range(B, rank, world)
    share = ((net(x[lo:hi]) - t[lo:hi]) ** 2).sum() / (B * 2)   # this rank's share of the global mean
    share.backward()
    tot32, tot64 = share.detach().clone(), share.detach().double().clone()
    D.allreduce_gradients(net, extra=[tot32, tot64])
    if rank == world - 1:
        torch.save(dict(grads={k: p.grad for k, p in net.named_parameters()}, tot32=tot32, tot64=tot64, rng=(lo, hi)), out)
    dist.destroy_process_group()


def test_uneven_shards_and_one_bucket_for_gradients_and_loss_shares(tmp_path):
    from dadmm_b200 import dist as D
    out = str(tmp_path / "last.pt")
    mp.spawn(_worker_uneven, args=(3, _free_port(), out), nprocs=3, join=True)
    got = torch.load(out)
    assert got["rng"] == (5, 7) and [D.shard_range(7, r, 3) for r in range(3)] == [(0, 3), (3, 5), (5, 7)]
    B, F = 7, 5
    g = torch.Generator().manual_seed(11)
    x, t = torch.randn(B, F, generator=g), torch.randn(B, 2, generator=g)
    torch.manual_seed(5)
    net = torch.nn.Sequential(torch.nn.Linear(F, 4), torch.nn.Tanh(), torch.nn.Linear(4, 2))
    loss = ((net(x) - t) ** 2).mean()
    loss.backward()
    ref = {k: p.grad for k, p in net.named_parameters()}
    assert got["grads"]["unused"] is None and set(got["grads"]) == set(ref) | {"unused"}
    for k, r in ref.items():
        assert torch.allclose(got["grads"][k], r, rtol=1e-5, atol=1e-7), k
    assert torch.allclose(got["tot32"], loss.detach(), rtol=1e-6) and got["tot64"].dtype == torch.float64
    assert abs(float(got["tot64"]) - float(loss.detach())) < 1e-6


def test_sharded_noise_is_the_slice_of_the_full_batch_draw():
    """Every rank draws the FULL batch under one seed and keeps its slice, so an N-rank run starts from the state a
    1-rank run starts from (three draws in the reference's order y, U, delta: unfolded_DLASSO.py:49-51)."""
    from dadmm_b200 import dist as D
    B, P, n, seed = 7, 3, 5, 123
    gen = torch.Generator().manual_seed(seed)
    full = [torch.randn((B, P, n, 1), generator=gen) * 1e-2 for _ in range(3)]
    parts = [D.sharded_noise(B, *D.shard_range(B, r, 3), P, n, "cpu", seed) for r in range(3)]
    for i in range(3):
        assert torch.equal(torch.cat([p[i] for p in parts]), full[i])
    assert all(t.is_contiguous() for p in parts for t in p)
