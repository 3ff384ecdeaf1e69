"""ctypes binding of libclasr_sm100.so (the C ABI declared in include/clasr_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this raises.  The
reference's convention is kept: non-zero status -> RuntimeError (reference
NeMo/nemo/collections/asr/parts/numba/rnnt_loss/rnnt.py:216-233).
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Dict, List, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libclasr_sm100.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "clasr_b200.h")

_lib = None

STATUS_SUCCESS = 0
ACT = {"relu": 0, "sigmoid": 1, "tanh": 2}
PREC = {"bf16": 0, "bf16x3": 1, "fp16x3": 2, "fp16m8": 3}

_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

_PROTOS: Dict[str, Tuple[object, List[object]]] = {
    "clasr_version": (_i, []),
    "clasr_last_error": (C.c_char_p, []),
    "clasr_launch_count": (_i64, []),
    "clasr_set_profiling": (None, [_i]),
    "clasr_profile_reset": (None, []),
    "clasr_profile_ms": (C.c_float, [C.c_char_p, _vp]),
    "clasr_cl_penalty_grad": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _f, _i, _vp, _vp]),
    "clasr_cl_penalty_avg": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "clasr_cl_fisher_accum": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "clasr_cl_mas_accum": (_i, [_vp, _vp, _i64, _vp]),
    "clasr_cl_scale_merge": (_i, [_vp, _vp, _i64, _f, _f, _i, _vp]),
    "clasr_cl_penalty_value_grad": (_i, [_vp, _vp, _vp, _i64, _f, _vp, _vp, _vp]),
    "clasr_cl_snapshot": (_i, [_vp, _vp, _i64, _vp]),
    "clasr_rnnt_workspace_bytes": (_sz, [_i, _i, _i]),
    "clasr_rnnt_loss_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _sz, _vp]),
    "clasr_rnnt_loss_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _vp, _vp, _vp, _sz, _vp]),
    "clasr_rnnt_export_lattice": (_i, [_vp, _sz, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "clasr_ctc_workspace_bytes": (_sz, [_i, _i, _i]),
    "clasr_ctc_loss_fwd": (_i, [_vp, _vp, _i64, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "clasr_ctc_loss_bwd": (_i, [_vp, _vp, _i64, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "clasr_log_softmax_fwd": (_i, [_vp, _vp, _i64, _i, _vp]),
    "clasr_log_softmax_bwd": (_i, [_vp, _vp, _vp, _i64, _i, _vp]),
    "clasr_gemm_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "clasr_gemm_nt": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "clasr_gemm_ex": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "clasr_debug_mma_rate": (_i, [_i, _i, _i, _vp, _vp]),
    "clasr_linear_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "clasr_linear_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "clasr_linear_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "clasr_transpose_last2": (_i, [_vp, _vp, C.c_int64, _i, _i, _vp]),
    "clasr_joint_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "clasr_joint_rnnt_fwd": (_i, [_vp] * 7 + [_i] * 8 + [_f, C.c_uint64, _f, _vp, _vp, _vp, _sz, _vp, _sz, _vp]),
    "clasr_joint_stash_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "clasr_joint_bwd_scratch_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "clasr_joint_sumsq_bwd": (_i, [_vp] * 7 + [_i] * 8 + [_f, C.c_uint64, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp, _sz, _vp]),
    "clasr_joint_rnnt_bwd": (_i, [_vp] * 7 + [_i] * 8 + [_f, C.c_uint64, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp, _sz, _vp]),
}


def declared_symbols() -> List[str]:
    """Every function include/clasr_b200.h declares (used by the ABI test)."""
    with open(HEADER_PATH) as fh:
        txt = fh.read()
    return sorted(set(re.findall(r"CLASR_API\s+[\w\s\*]+?\b(clasr_\w+)\s*\(", txt)))


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.environ.get("CLASR_LIB", LIB_PATH)   # developer override (A/B builds); the product path is LIB_PATH
        if not os.path.exists(path):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C indic_cl_asr_b200/csrc).  There is no CPU / PyTorch fallback."
            )
        l = C.CDLL(path)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def last_error() -> str:
    return lib().clasr_last_error().decode()


def check(status: int, what: str = "") -> None:
    if status != STATUS_SUCCESS:
        raise RuntimeError(f"clasr {what} failed with status {status}: {last_error()}")


def stream_ptr(device=None) -> int:
    """cudaStream_t of torch's CURRENT stream (reference: rnnt.py:173-176)."""
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} must be a CUDA tensor: indic_cl_asr_b200 has no CPU path (the CPU implementation lives only in "
            "the test oracle)"
        )


def launch_count() -> int:
    return int(lib().clasr_launch_count())


def profile_ms_count(name: str):
    """(mean ms per launch, number of launches) recorded under `name` since the last reset."""
    n = C.c_int(0)
    ms = float(lib().clasr_profile_ms(name.encode(), C.byref(n)))
    return ms, int(n.value)


def profile_ms(name: str) -> float:
    """Mean ms of the library kernel recorded under `name` since the last reset (-1 if none)."""
    return float(lib().clasr_profile_ms(name.encode(), None))
