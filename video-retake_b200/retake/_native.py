"""ctypes binding of ``librtk_b200.so`` (``include/rtk_b200.h``).

The library is the product: there is no fallback.  If it is missing, or a tensor is not a CUDA
tensor, the call raises - nothing here ever computes on the host or through stock torch ops.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RTK_B200_LIB", os.path.join(os.path.dirname(_HERE), "lib", "librtk_b200.so"))

_lib = None
ABI_VERSION = 3          # RTK_ABI_VERSION of include/rtk_b200.h


class RtkError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RtkError(
            f"{LIB_PATH} not found: build it with `python video-retake_b200/build.py` "
            "(nvcc, sm_100a). There is no CPU or stock-PyTorch fallback for this path."
        )
    L = C.CDLL(LIB_PATH)
    p, i64, i32, f32, sz = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_size_t
    sig = {
        "rtk_version": ([], C.c_int),
        "rtk_build_id": ([], C.c_char_p),
        "rtk_debug_key_elision": ([i32], C.c_int),
        "rtk_error_string": ([C.c_int], C.c_char_p),
        "rtk_launch_count": ([], i64),
        "rtk_dpselect_dis": ([p, i64, i64, i64, i32, p, p], C.c_int),
        "rtk_dpselect_select": ([p, i64, i64, i64, i32, p, p, p], C.c_int),
        "rtk_dpselect_gather": ([p, i64, i64, i64, p, i64, i32, p, p], C.c_int),
        "rtk_gather_rows": ([p, i64, p, i64, p, p], C.c_int),
        "rtk_dpselect_keyframe": ([p, i64, i64, i64, i64, i32, p, p, p, p, p], C.c_int),
        "rtk_dpselect_gather_owned": ([p, i64, i64, i64, i64, i64, i64, p, i64, i32, p, p], C.c_int),
        "rtk_mallm_workspace_bytes": ([i64, i64, i64, i32], sz),
        "rtk_mallm_compress": ([p, p, i64, i64, i64, i64, i32, i32, p, p, p, sz, p], C.c_int),
        "rtk_pivot_rope": ([p, i64, i64, i64, i64, i64, p, p, i32, p, f32, i32, p, i64, i64, p], C.c_int),
        "rtk_pivot_score_workspace_bytes": ([i64, i64], sz),
        "rtk_pivot_score": ([p, i64, i64, i64, p, i64, i64, i64, i64, i64, p, p, sz, p], C.c_int),
        "rtk_pivot_select": ([p, i64, i64, p, i64, p, p, p], C.c_int),
        "rtk_pivot_compact": ([p, p, i64, i64, i64, i64, i64, p, i64, p, p, i64, p, i32, p, i32, p], C.c_int),
        "rtk_pivot_rope_tables": ([p, i32, i64, i64, p, p, f32, p, p, p], C.c_int),
        "rtk_pivot_update_workspace_bytes": ([i64, i64, i64, i64], sz),
        "rtk_pivot_update": ([p, p], C.c_int),
        "rtk_pivot_update_batch_workspace_bytes": ([i64, i64, i64, i64, i64], sz),
        "rtk_pivot_update_batch": ([p, i64, p, sz, p], C.c_int),
        "rtk_kv_block_copy": ([i32, p, p, p, p, p, p, p, p, i64, p], C.c_int),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)            # AttributeError here == header and library disagree
        fn.argtypes, fn.restype = args, res
    if L.rtk_version() != ABI_VERSION:
        raise RtkError(f"ABI mismatch: library reports version {L.rtk_version()}, binding expects {ABI_VERSION}")
    _check_build_id(L)
    _lib = L
    return L


def _source_id():
    """id of the sources lying next to this package (None when the tree ships without them)"""
    import importlib.util
    path = os.path.join(os.path.dirname(_HERE), "build.py")
    if not os.path.exists(path) or not os.path.isdir(os.path.join(os.path.dirname(_HERE), "csrc")):
        return None
    spec = importlib.util.spec_from_file_location("rtk_build_for_id", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.source_id()


def _check_build_id(L) -> None:
    """the in-tree library must have been built from the in-tree sources (the build used to be gated on file
    times, so nothing proved which sources a record came from).  RTK_B200_LIB (an explicitly chosen library, e.g. an A/B
    variant) and RTK_ALLOW_STALE_LIB=1 skip the check."""
    if "RTK_B200_LIB" in os.environ or os.environ.get("RTK_ALLOW_STALE_LIB") == "1":
        return
    want = _source_id()
    got = L.rtk_build_id().decode()
    if want is not None and got != want:
        raise RtkError(f"{LIB_PATH} was built from other sources (library id {got}, sources {want}): "
                       "run `python video-retake_b200/build.py`")


def build_id() -> str:
    return lib().rtk_build_id().decode()


EXPORTS = ("rtk_version", "rtk_build_id", "rtk_debug_key_elision", "rtk_error_string", "rtk_launch_count", "rtk_dpselect_dis", "rtk_dpselect_select",
           "rtk_dpselect_gather", "rtk_dpselect_keyframe", "rtk_gather_rows", "rtk_dpselect_gather_owned", "rtk_mallm_workspace_bytes", "rtk_mallm_compress", "rtk_pivot_rope", "rtk_pivot_score_workspace_bytes", "rtk_pivot_score",
           "rtk_pivot_select", "rtk_pivot_compact", "rtk_pivot_rope_tables", "rtk_pivot_update_workspace_bytes",
           "rtk_pivot_update", "rtk_pivot_update_batch_workspace_bytes", "rtk_pivot_update_batch", "rtk_kv_block_copy")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RtkError(f"{what} failed: {lib().rtk_error_string(rc).decode()} (code {rc})")


def require_cuda(t: torch.Tensor, name: str, dtype=None) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RtkError(f"{name} must be a CUDA tensor: this path has no CPU implementation")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def launch_count() -> int:
    return int(lib().rtk_launch_count())
