"""ctypes binding of ``libdadmm_sm100.so`` (C ABI in ``include/dadmm.h``).

The library is the product: there is NO CPU or PyTorch fallback.  Importing this module without
the built library raises; calling a compute entry point with non-CUDA tensors raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("DADMM_LIB", os.path.join(_PKG_DIR, "libdadmm_sm100.so"))

ABI_VERSION = 7
F32, F64 = 0, 1
ALGO_AUTO, ALGO_SIMT, ALGO_TC_3XTF32, ALGO_TC_3XF16, ALGO_TC_F16X1 = 0, 1, 2, 3, 4
ALGOS = {"auto": ALGO_AUTO, "simt": ALGO_SIMT, "tc": ALGO_TC_3XTF32, "tc_3xtf32": ALGO_TC_3XTF32, "tf32": ALGO_TC_3XTF32,
         "f16": ALGO_TC_3XF16, "tc_3xf16": ALGO_TC_3XF16,
         "fast": ALGO_TC_F16X1, "f16x1": ALGO_TC_F16X1}      # "fast": flagged reduced-precision mode (1e-2 class)
FLAG_Y, FLAG_U, FLAG_GRAD, FLAG_YNEXT = 1, 2, 4, 8


class Graph(C.Structure):
    _fields_ = [("n_graphs", C.c_int32), ("P", C.c_int32), ("ev_ptr", C.c_void_p), ("ev_idx", C.c_void_p),
                ("deg", C.c_void_p), ("graph_id", C.c_void_p), ("adj_ptr", C.c_void_p), ("adj_idx", C.c_void_p),
                ("max_events", C.c_int32), ("max_adj", C.c_int32)]


class Clamps(C.Structure):
    _fields_ = [("G", C.c_double), ("V", C.c_double), ("D", C.c_double), ("Uc", C.c_double)]


class Hyp(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("stride_b", C.c_int64), ("stride_p", C.c_int64), ("stride_c", C.c_int64)]


class Factor(C.Structure):
    """W_p = F2_p F1_p, F1 [P,m,n], F2 [P,n,m] (``dadmm_factor``)."""
    _fields_ = [("m", C.c_int32), ("F1", C.c_void_p), ("F2", C.c_void_p), ("rhs", C.c_void_p)]


class LossSums(C.Structure):
    """Label-free per-iteration sums written by ``dadmm_unfolded_fwd`` (``dadmm_loss_sums``)."""
    _fields_ = [("agent_sum", C.c_void_p), ("sumsq", C.c_void_p), ("valid", C.POINTER(C.c_int32))]


class OpSplit(C.Structure):
    """Persistent home of the operator's tensor-core operand copies (``dadmm_op_split``)."""
    _fields_ = [("buf", C.c_void_p), ("bytes", C.c_size_t), ("ready", C.c_int32)]


class DadmmError(RuntimeError):
    pass


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` (nvcc, sm_100a). "
            "There is no CPU fallback for the D-ADMM hot path.")
    lib = C.CDLL(LIB_PATH)
    i32, i64, vp, dbl, sz = C.c_int, C.c_int64, C.c_void_p, C.c_double, C.c_size_t
    GP, CP, HP, FP, SP = C.POINTER(Graph), C.POINTER(Clamps), C.POINTER(Hyp), C.POINTER(Factor), C.POINTER(LossSums)
    OP = C.POINTER(OpSplit)
    sigs = {
        "dadmm_abi_version": (i32, []),
        "dadmm_last_error": (C.c_char_p, []),
        "dadmm_device_check": (i32, []),
        "dadmm_launch_count": (i64, []),
        "dadmm_set_pdl": (i32, [i32]),
        "dadmm_set_consensus_order": (i32, [i32]),
        "dadmm_profile_enable": (i32, [i32]),
        "dadmm_profile_read": (i32, [C.POINTER(dbl), C.POINTER(i64)]),
        "dadmm_contract": (i32, [i32, i32, i32, i32, i32, i32, vp, i64, i64, i64, vp, i64, i64, i64, vp, i64, i64, i64,
                                 i32, vp, sz, vp]),
        "dadmm_contract_prepared": (i32, [i32, i32, i32, i32, i32, i32, vp, i64, i64, i64, vp, i64, i64, i64, vp, i64, i64, i64,
                                          i32, vp, sz, i32, vp]),
        "dadmm_contract_ws_bytes": (sz, [i32, i32, i32, i32, i32, i32]),
        "dadmm_contract_uses_tensor_cores": (i32, [i32, i32, i32, i32, i32, i32]),
        "dadmm_step_fwd": (i32, [i32, i32, i32, i32, GP, CP, HP, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "dadmm_step_bwd": (i32, [i32, i32, i32, i32, GP, CP, HP, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, dbl,
                                 vp, vp, vp, vp, vp, vp]),
        "dadmm_partials_elems": (sz, [i32, i32, i32, i32]),
        "dadmm_reduce_hyp": (i32, [i32, i32, i32, i32, vp, i32, vp, i64, i64, i64, i32, vp]),
        "dadmm_unfolded_fwd": (i32, [i32, i32, i32, i32, i32, i32, GP, CP, vp, vp, FP, vp, vp, vp, vp, vp, vp, vp, vp, sz,
                                     vp, SP, OP, vp]),
        "dadmm_unfolded_bwd": (i32, [i32, i32, i32, i32, i32, i32, GP, CP, vp, vp, FP, vp, vp, vp, vp, vp, vp, vp, vp,
                                     C.POINTER(dbl), vp, vp, vp, sz, OP, vp]),
        "dadmm_unfolded_op_split_bytes": (sz, [i32, i32, i32, i32, i32, i32]),
        "dadmm_unfolded_ws_bytes": (sz, [i32, i32, i32, i32, i32, i32, i32, i32]),
        "dadmm_unfolded_uses_factor": (i32, [i32, i32, i32, i32, i32, i32]),
        "dadmm_loss_fwd": (i32, [i32, i32, i32, i32, i32, i64, vp, vp, vp, vp, sz, vp]),
        "dadmm_loss_bwd": (i32, [i32, i32, i32, i32, i32, vp, vp, C.POINTER(dbl), vp, vp]),
        "dadmm_loss_ws_bytes": (sz, [i32, i32, i32, i32, i32]),
        "dadmm_loss_from_sums": (i32, [i32, i32, i32, i32, i32, i32, i64, vp, vp, vp, vp, vp, vp, sz, vp]),
        "dadmm_gcn_epilogue_fwd": (i32, [i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, i32, dbl, dbl, vp, vp, vp, vp, vp, vp]),
        "dadmm_gcn_epilogue_bwd": (i32, [i32, i32, i32, vp, vp, vp, vp, vp, i32, dbl, dbl, vp, vp, vp, vp, vp, vp, vp]),
        "dadmm_gcn_partial_rows": (i32, [i32, i32]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.dadmm_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version {lib.dadmm_abi_version()} != {ABI_VERSION}")
    return lib, tuple(sigs)


lib, EXPORTED = _load()


def check(rc: int, what: str):
    if rc != 0:
        raise DadmmError(f"{what} failed (rc={rc}): {lib.dadmm_last_error().decode(errors='replace')}")


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float64:
        return F64
    raise TypeError(f"libdadmm_sm100 computes in float32/float64, got {t.dtype}")


def require_cuda(*tensors: Optional[torch.Tensor]):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise DadmmError("the D-ADMM hot path runs on a B200 (sm_100a) only: got a CPU tensor and there is "
                             "no CPU fallback -- move the inputs to cuda")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise DadmmError(f"tensors on different devices: {dev} vs {t.device}")
    return dev


_checked_devices = set()


def device_guard(dev: torch.device):
    """Make ``dev`` current (the library launches on the current device) and verify sm_100 once."""
    guard = torch.cuda.device(dev)
    guard.__enter__()
    if dev.index not in _checked_devices:
        try:
            check(lib.dadmm_device_check(), "dadmm_device_check")
        except Exception:
            guard.__exit__(None, None, None)
            raise
        _checked_devices.add(dev.index)
    return guard


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(dev) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def set_pdl(on: bool) -> bool:
    """Programmatic dependent launch of the K-loop chain on/off (``dadmm_set_pdl``); returns the previous setting."""
    return bool(lib.dadmm_set_pdl(int(bool(on))))


def set_consensus_order(exact: bool) -> bool:
    """2 L y of the fused forward levels in the reference's event order (bit-identical delta) or with every neighbour
    difference taken once and doubled (default; ``dadmm_set_consensus_order``); returns the previous setting."""
    return bool(lib.dadmm_set_consensus_order(int(bool(exact))))


def launch_count() -> int:
    return int(lib.dadmm_launch_count())


PROF_KINDS = ("contract_simt", "contract_tc", "step_fwd", "step_bwd", "reduce_hyp", "loss", "split", "contract_stage1")


def profile_enable(on: bool = True):
    check(lib.dadmm_profile_enable(int(on)), "dadmm_profile_enable")


def profile_read():
    """{kind: (total_ms, launches)} accumulated since profile_enable(True)."""
    ms, cnt = (C.c_double * 8)(), (C.c_int64 * 8)()
    check(lib.dadmm_profile_read(ms, cnt), "dadmm_profile_read")
    return {k: (ms[i], int(cnt[i])) for i, k in enumerate(PROF_KINDS)}
