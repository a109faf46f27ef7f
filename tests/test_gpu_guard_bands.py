"""Out-of-bounds WRITE check of the CUDA path without a sanitizer (compute-sanitizer is not available on the GPU pool).

Every device buffer the host side allocates while these tests run -- outputs, saved tensors, workspaces, partial-sum rows
-- is carved out of the middle of a larger allocation whose first and last 4 KiB hold a canary byte.  After the run
(forward, loss, backward) every canary must be intact: a kernel that stores one element past the end (or before the start)
of anything it was handed fails the test.  The shapes take the routes the bench workloads take (two-stage tcgen05
contraction + LEAN levels), the ragged ones (n = 500: generic levels, FP32-FMA contraction), the float64 instantiation,
the inference sweep and the model-#3 step with the graph-convolution kernels.  Results are compared with an unguarded run
of the same call (bit-identical), so the guard itself cannot mask a difference.
"""
import argparse
import math

import pytest
import torch

from helpers import random_problem

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PAD = 4096                      # bytes each side; a multiple of every alignment the kernels ask for (TMA: 128 B)
CANARY = 0xA5


class GuardBands:
    """Context manager: ``torch.empty`` / ``empty_like`` / ``zeros`` / ``zeros_like`` on CUDA return the middle of a padded
    allocation; ``check()`` verifies the padding."""

    def __init__(self):
        self.records = []
        self._orig = {}

    def _alloc(self, shape, dtype, device, zero):
        shape = tuple(int(s) for s in shape)
        es = torch.empty((), dtype=dtype).element_size()
        n = math.prod(shape)
        pad = PAD // es
        flat = self._orig["empty"](n + 2 * pad, dtype=dtype, device=device)
        raw = flat.view(torch.uint8)
        raw[:PAD] = CANARY
        raw[PAD + n * es:] = CANARY
        self.records.append((raw, n * es, shape, dtype))
        mid = flat[pad:pad + n]
        if zero:
            mid.zero_()
        return mid.view(shape)

    @staticmethod
    def _shape(size):
        if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
            return tuple(size[0])
        return tuple(size)

    def _wrap(self, name, zero):
        orig = self._orig[name]

        def fn(*size, dtype=None, device=None, **kw):
            dev = torch.device(device) if device is not None else None
            plain = dev is None or dev.type != "cuda" or "size" in kw or any(kw.get(k) is not None and kw.get(k) is not False for k in ("pin_memory", "out", "names")) \
                or kw.get("memory_format", torch.contiguous_format) != torch.contiguous_format or kw.get("layout", torch.strided) != torch.strided
            if plain:
                return orig(*size, dtype=dtype, device=device, **kw)
            t = self._alloc(self._shape(size), dtype or torch.get_default_dtype(), dev, zero)
            return t.requires_grad_(True) if kw.get("requires_grad") else t
        return fn

    def _wrap_like(self, name, zero):
        orig = self._orig[name]

        def fn(t, dtype=None, device=None, **kw):
            dev = torch.device(device) if device is not None else t.device
            if dev.type != "cuda" or not t.is_contiguous() or kw.get("memory_format", torch.preserve_format) not in (torch.preserve_format, torch.contiguous_format) \
                    or t.layout != torch.strided:
                return orig(t, dtype=dtype, device=device, **kw)
            return self._alloc(t.shape, dtype or t.dtype, dev, zero)
        return fn

    def __enter__(self):
        for name in ("empty", "zeros", "empty_like", "zeros_like"):
            self._orig[name] = getattr(torch, name)
        torch.empty, torch.zeros = self._wrap("empty", False), self._wrap("zeros", True)
        torch.empty_like, torch.zeros_like = self._wrap_like("empty_like", False), self._wrap_like("zeros_like", True)
        return self

    def __exit__(self, *exc):
        for name, fn in self._orig.items():
            setattr(torch, name, fn)
        return False

    def check(self):
        torch.cuda.synchronize()
        assert self.records, "the guarded run allocated nothing through torch.empty/zeros: the check would be vacuous"
        bad = []
        for raw, nbytes, shape, dtype in self.records:
            lo, hi = raw[:PAD], raw[PAD + nbytes:]
            if not (bool((lo == CANARY).all()) and bool((hi == CANARY).all())):
                before = int((lo != CANARY).sum())
                after = int((hi != CANARY).sum())
                bad.append((shape, dtype, before, after))
        assert not bad, f"out-of-bounds writes next to {len(bad)} of {len(self.records)} buffers (shape, dtype, bytes before, bytes after): {bad[:8]}"
        return len(self.records)


def _args(P, n, m, K, mode="diff", **kw):
    a = argparse.Namespace(m=m, n=n, P=P, GHN_iter_num=K, DADMM_mode=mode, alpha_max=0.1, tau_max=0.99, rho_max=0.99,
                           eta_max=0.99, max_penalty_threshold=0.8, penalty_reduction_factor=0.95, batch_size=1, snr=4,
                           GHyp_hidden=16.0)
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def _model1_step(P, n, m, B, K, dtype, per_sample, grad=True, seed=0):
    import unfolded_DLASSO
    import gnn_dlasso_utils
    from dadmm_b200 import functional as DF
    prob = random_problem(P, n, m, B, K, seed=seed, per_sample_graphs=per_sample, dtype=dtype)
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        DF.clear_caches()
        model = unfolded_DLASSO.DLASSO_unfolded(prob["A"].to(DEV), _args(P, n, m, K)).to(DEV)
        model.load_state_dict({"seq_hyp.param": prob["param"]})
        noise = tuple(prob[k].to(DEV) for k in ("y0", "U0", "d0"))
        b, label = prob["b"].to(DEV), prob["label"].to(DEV)
        if not grad:
            with torch.no_grad():
                Y, _ = model(b, prob["graphs"], noise=noise)
            return Y.clone(), None
        Y, _ = model(b, prob["graphs"], noise=noise)
        lm, lf = gnn_dlasso_utils.compute_loss(Y, label)
        lf.backward()
        return Y.detach().clone(), model.seq_hyp.param.grad.clone()
    finally:
        torch.set_default_dtype(old)


MODEL1_SHAPES = [
    # P, n, m, B, K, dtype, per-sample graphs, backward
    pytest.param(4, 512, 192, 256, 3, torch.float32, True, True, id="two_stage_tcgen05_lean"),
    pytest.param(6, 256, 64, 160, 3, torch.float32, True, True, id="single_stage_tcgen05_lean_ragged_batch"),
    pytest.param(5, 500, 100, 32, 4, torch.float32, False, True, id="fma_contraction_generic_levels_shared_graph"),
    pytest.param(7, 130, 40, 9, 3, torch.float32, True, True, id="odd_everything"),
    pytest.param(5, 96, 32, 12, 3, torch.float64, True, True, id="float64"),
    pytest.param(12, 384, 128, 130, 4, torch.float32, True, False, id="inference_sweep"),
]


@pytest.mark.parametrize("P,n,m,B,K,dtype,per_sample,grad", MODEL1_SHAPES)
def test_model1_step_writes_inside_its_buffers(P, n, m, B, K, dtype, per_sample, grad):
    Y_ref, g_ref = _model1_step(P, n, m, B, K, dtype, per_sample, grad)
    with GuardBands() as gb:
        Y, g = _model1_step(P, n, m, B, K, dtype, per_sample, grad)
        count = gb.check()
    print(f"guarded buffers: {count}")
    assert torch.equal(Y, Y_ref)
    if grad:
        assert torch.equal(g, g_ref)


def test_exact_consensus_order_writes_inside_its_buffers():
    from dadmm_b200 import _lib
    _lib.set_consensus_order(True)
    try:
        Y_ref, g_ref = _model1_step(4, 512, 192, 256, 3, torch.float32, True)
        with GuardBands() as gb:
            Y, g = _model1_step(4, 512, 192, 256, 3, torch.float32, True)
            gb.check()
    finally:
        _lib.set_consensus_order(False)
    assert torch.equal(Y, Y_ref) and torch.equal(g, g_ref)


def _model3_step(B, P, n, m, K, hidden, seed=0):
    import gnn_dlasso_models_progressive as M
    import gnn_dlasso_utils
    from dadmm_b200 import functional as DF
    prob = random_problem(P, n, m, B, K, seed=seed, per_sample_graphs=True)
    DF.clear_caches()
    torch.manual_seed(5)
    model = M.DLASSO_GNNHyp3_Progressive(prob["A"].to(DEV), _args(P, n, m, K, GHyp_hidden=float(hidden))).to(DEV).train()
    model.encoder.dropout.p = 0.0                      # same activations in both runs
    for mod in model.decoder:
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    Y, _ = model(prob["b"].to(DEV), prob["graphs"], noise=tuple(prob[k].to(DEV) for k in ("y0", "U0", "d0")))
    lm, lf = gnn_dlasso_utils.compute_loss(Y, prob["label"].to(DEV))
    lf.backward()
    return Y.detach().clone(), model.fc.weight.grad.clone(), model.encoder.conv1.lin.weight.grad.clone()


def test_model3_step_writes_inside_its_buffers():
    ref = _model3_step(48, 5, 64, 24, 3, 24)
    with GuardBands() as gb:
        got = _model3_step(48, 5, 64, 24, 3, 24)
        gb.check()
    for a, r in zip(got, ref):
        assert torch.equal(a, r)
