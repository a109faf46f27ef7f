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


def test_two_rank_sharding_matches_single_process(tmp_path):
    from oracle import dadmm_oracle as O
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, 29533, out), nprocs=2, join=True)
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
