#!/usr/bin/env python
"""One training step (forward K iterations + loss + backward) of the bench workload with a reduced K, for ncu:

    python tools/profile_step.py [--workload cfg4] [--K 2] [--batch 4096] [--algo auto]
    ncu --set full --clock-control none --import-source on -k regex:contract_tc -c 2 -o gpurun_out/prof python tools/profile_step.py
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg4")
    ap.add_argument("--K", type=int, default=2)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--algo", default="auto")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--inference", action="store_true", help="forward only under no_grad (config-5 style sweep)")
    o = ap.parse_args()
    import unfolded_DLASSO
    import gnn_dlasso_utils
    w = dict(bench.WORKLOADS[o.workload])
    w["K"] = o.K
    B = o.batch or w["B"]
    w["B"] = B
    dev = torch.device("cuda:0")
    args, A, label, graphs, param = bench.make_problem(w, B)
    A, label = A.to(dev), label.to(dev)
    b = torch.stack([A[0, p] @ label for p in range(w["P"])], dim=1).contiguous()
    model = unfolded_DLASSO.DLASSO_unfolded(A, args).to(dev)
    model.contract_algo = o.algo
    with torch.no_grad():
        model.seq_hyp.param.copy_(param)
    import time
    for it in range(o.steps + 1):
        if it == 1:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        if o.inference:
            with torch.no_grad():
                Y, _ = model(b, graphs)
                lm, lf = gnn_dlasso_utils.compute_loss(Y, label, check_finite=False)
        else:
            Y, _ = model(b, graphs)
            lm, lf = gnn_dlasso_utils.compute_loss(Y, label, check_finite=False)
            model.zero_grad()
            lf.backward()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / max(o.steps, 1)
    print("ok", float(lf.detach()), f"{1e3*dt:.2f} ms/step", f"{o.K*B/dt:.0f} iter*problems/s", f"peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")


if __name__ == "__main__":
    main()
