"""N ranks (torchrun, one per GPU) against one process on the same batch: every rank solves its contiguous shard with its
slice of the FULL batch's initial noise (``dadmm_b200.dist.sharded_noise`` -> ``forward(..., noise=...)``), scales its loss by
the global batch and all-reduces loss and gradients; rank 0 then solves the whole batch alone and compares.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/probes/rank_parity.py

Prints one line starting with RANK_PARITY; exit code 0 when the N-rank loss, gradient and iterates agree with the single
process to 1e-5 / 1e-5 / 2e-6 (contracting regime: not bit for bit -- the tensor-core contraction scales its fp16 operand
pairs by a power of two taken from the max |.| of the tensor it is handed, i.e. of the shard)."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import random_problem, rel_l2  # noqa: E402  (also puts the package on sys.path)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import unfolded_DLASSO
    import gnn_dlasso_utils
    from dadmm_b200 import dist as D
    P, n, m, B, K = 6, 512, 160, 512, 8
    pr = random_problem(P, n, m, B, K, seed=9, a_scale=0.1)
    gen = torch.Generator().manual_seed(3)
    param = torch.randn((K, P, 4), generator=gen) * 0.05
    param[0] += torch.tensor([-1.0, -2.0, -4.0, -4.0])          # contracting regime (tests/test_gpu_baseline_shapes.py)
    args = argparse.Namespace(m=m, n=n, P=P, GHN_iter_num=K, DADMM_mode="diff", alpha_max=0.1, tau_max=0.99, rho_max=0.99,
                              eta_max=0.99, max_penalty_threshold=0.8, penalty_reduction_factor=0.95, batch_size=B, snr=4)

    def solve(lo, hi):
        model = unfolded_DLASSO.DLASSO_unfolded(pr["A"].to(dev), args).to(dev)
        with torch.no_grad():
            model.seq_hyp.param.copy_(param)
        noise = D.sharded_noise(B, lo, hi, P, n, dev, seed=77)
        Y, _ = model(pr["b"][lo:hi].to(dev), pr["graphs"][lo:hi], noise=noise)
        _, lf = gnn_dlasso_utils.compute_loss(Y, pr["label"][lo:hi].to(dev), global_batch=B)
        lf.backward()
        return model, Y.detach(), lf.detach().clone()

    lo, hi = D.shard_range(B, rank, world)
    model, Y, loss = solve(lo, hi)
    loss_share = loss - 1e-8 * (rank != 0)                  # compute_loss adds its 1e-8 on every rank
    D.allreduce_gradients(model, extra=[loss_share])
    ok = True
    if rank == 0:
        full, Yf, lf = solve(0, B)
        eY = rel_l2(Y.cpu(), Yf[:, lo:hi].cpu())
        eg = rel_l2(model.seq_hyp.param.grad.cpu(), full.seq_hyp.param.grad.cpu())
        el = abs(float(loss_share) - float(lf)) / abs(float(lf))
        ok = eY < 2e-6 and eg < 1e-5 and el < 1e-5
        print(f"RANK_PARITY world={world} rel_l2(Y shard vs single)={eY:.2e} rel_l2(grad)={eg:.2e} rel(loss)={el:.2e} "
              f"loss={float(loss_share):.7f} single={float(lf):.7f} {'OK' if ok else 'MISMATCH'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
