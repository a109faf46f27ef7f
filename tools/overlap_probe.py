#!/usr/bin/env python
"""Would two independent half-batch pipelines on two streams overlap the tensor-bound contractions of one half with the
HBM-bound level kernels of the other?  Aggregate throughput of 2 x (B/2) on two streams / threads vs 1 x B.
    python tools/overlap_probe.py [--K 25] [--steps 3] [--prio]"""
import argparse, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
ap = argparse.ArgumentParser(); ap.add_argument("--K", type=int, default=25); ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--parts", type=int, default=2)
o = ap.parse_args()
import unfolded_DLASSO, gnn_dlasso_utils
w = dict(bench.WORKLOADS["cfg4"]); w["K"] = o.K
dev = torch.device("cuda:0")
args, A, label, graphs, param = bench.make_problem(w, w["B"])
A, label = A.to(dev), label.to(dev)
b = torch.stack([A[0, p] @ label for p in range(w["P"])], dim=1).contiguous()

def make(lo, hi):
    m = unfolded_DLASSO.DLASSO_unfolded(A, args).to(dev)
    with torch.no_grad(): m.seq_hyp.param.copy_(param)
    return m, b[lo:hi].contiguous(), label[lo:hi].contiguous(), graphs[lo:hi]

def step(m, bb, ll, gg):
    Y, _ = m(bb, gg); lm, lf = gnn_dlasso_utils.compute_loss(Y, ll, check_finite=False, global_batch=w["B"]); m.zero_grad(); lf.backward()

full = make(0, w["B"])
for _ in range(2): step(*full)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(o.steps): step(*full)
torch.cuda.synchronize(); t_full = (time.perf_counter() - t0) / o.steps
print(f"1 x {w['B']}: {1e3*t_full:.1f} ms/step")
del full; torch.cuda.empty_cache()

n = o.parts
parts = [make(i * w["B"] // n, (i + 1) * w["B"] // n) for i in range(n)]
streams = [torch.cuda.Stream(dev) for _ in range(n)]
def worker(i, steps):
    with torch.cuda.stream(streams[i]):
        for _ in range(steps): step(*parts[i])
def run(steps):
    th = [threading.Thread(target=worker, args=(i, steps)) for i in range(n)]
    [t.start() for t in th]; [t.join() for t in th]; torch.cuda.synchronize()
run(2)
t0 = time.perf_counter(); run(o.steps); t_par = (time.perf_counter() - t0) / o.steps
print(f"{n} x {w['B']//n} on {n} streams: {1e3*t_par:.1f} ms/step  -> {t_full/t_par:.3f}x")
