"""Whole-step CUDA-graph capture for the host-bound shapes.

BASELINE configs[0] (P=5, n=500, batch 32) is a chain of ~100 library launches plus the optimizer's and the loss's few
PyTorch kernels: under 1 ms of GPU work that takes 2.2 ms per step when it is launched from Python (round-2 measurement);
the model-#3 step (configs[1]) is bound by the ~4700 launches of its hypernetwork in the same way.  The library never
allocates, takes its stream as an argument and reads nothing back, so a training step -- ``model(b, graph_list)``,
``compute_loss``, ``loss.backward()``, ``optimizer.step()`` -- can be captured once and replayed as ONE graph launch.

What the capture needs from the modules (all of it opt-in; the eager path is unchanged):
  * no host read inside the step: ``model.check_finite = "deferred"`` keeps the kernels' non-finite flags in a sticky
    device buffer instead of reading them after every forward (``model.nonfinite_seen()`` reads it when the caller wants
    to know -- e.g. once per epoch); ``compute_loss(..., check_finite=False)``;
  * static inputs: ``GraphedStep`` owns the input buffers and copies each step's batch into them;
  * the same ``graph_list`` (or ``BatchGraph``) object for every replay -- its device arrays are part of the graph;
  * an optimizer built with ``capturable=True``.
The reference's reset / skip semantics for non-finite values (unfolded_DLASSO.py:55-61,84-86,102-104) need a host decision
per iteration and therefore cannot live inside a graph: a replay that raises the sticky flag has run the plain recurrence
on non-finite data -- rerun that batch eagerly (``check_finite=True``) to get the guarded result.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class GraphedStep:
    """``fn(*inputs) -> tensor or tuple of tensors`` captured into one CUDA graph.

    ``fn`` is called ``warmup`` times eagerly on a side stream (lazy initialisation, operator caches, autotuning happen
    there), then once under capture.  ``__call__`` copies the new inputs into the captured buffers, replays, and returns
    the captured output tensors (overwritten by the next replay: clone what must survive)."""

    def __init__(self, fn: Callable, example_inputs: Sequence[torch.Tensor], warmup: int = 3):
        self.fn = fn
        self.static_in = [t.detach().clone() for t in example_inputs]
        dev = self.static_in[0].device
        self.stream = torch.cuda.Stream(dev)
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.stream):
            for _ in range(max(warmup, 1)):
                fn(*self.static_in)
        torch.cuda.current_stream(dev).wait_stream(self.stream)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        # warm-up and capture share one stream: per-stream state (the persistent operator split) made during warm-up is
        # what the captured launches read
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.static_out = fn(*self.static_in)
        self.replays = 0

    def __call__(self, *inputs: torch.Tensor):
        for dst, src in zip(self.static_in, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        return self.static_out
