"""DPSelect on B200: drop-in for ``retake/visual_compression.py`` of SCZwangxiao/video-ReTaKe.

``memory_bank_compress_keyframe`` keeps the reference's name, arguments, return values and error
behaviour (``visual_compression.py:86-177``); its body is three sm_100a kernel launches through the
C ABI (``include/rtk_b200.h``) on the current CUDA stream, with no host synchronisation:

    rtk_dpselect_dis     adjacent-frame cosine distance            (reference lines 98-106)
    rtk_dpselect_select  peaks, +2 priority, top-t, ascending sort (lines 108-135 / 141-169,175)
    rtk_dpselect_gather  compaction of the surviving tokens        (lines 138 / 173)

The alternative compressors of the reference (``memory_bank_compress_MALLM*``, lines 5-83; ``compression_method:
MA-LLM / MA-LLM-hard``) run on ``rtk_mallm_compress``: the single-round functions keep the reference's signatures,
``mallm_compress`` is the caller's whole ``while T > tgt`` loop (``qwen2_vl.py:402-409``) in one library call.
"""
from __future__ import annotations

import torch

from . import _native as N

__all__ = ["memory_bank_compress_keyframe", "dpselect_distance", "dpselect_select", "dpselect_gather",
           "memory_bank_compress_MALLM", "memory_bank_compress_MALLM_hard", "mallm_compress"]


def dpselect_distance(x: torch.Tensor, halo: bool = False) -> torch.Tensor:
    """``dis[T, N]`` fp32 for ``x[T, N, C]`` bf16.  With ``halo`` the first frame belongs to the previous
    frame range (multi-GPU split) and the result has T-1 rows."""
    N.require_cuda(x, "memory_bank", torch.bfloat16)
    if x.dim() != 3:
        raise ValueError("expected [T, N, C]")
    x = x.contiguous()
    T, Np, Cc = x.shape
    rows = T - 1 if halo else T
    dis = torch.empty((rows, Np), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        N.check(N.lib().rtk_dpselect_dis(x.data_ptr(), T, Np, Cc, int(halo), dis.data_ptr(), N.stream_ptr(x.device)),
                "rtk_dpselect_dis")
    return dis


def dpselect_select(dis: torch.Tensor, t: int, sync: bool):
    """Kept frame indices (int32 ``[t, N]`` or ``[t]``, ascending) and the flat key-patch mask (bool ``[t*N]``)."""
    N.require_cuda(dis, "dis", torch.float32)
    dis = dis.contiguous()
    T, Np = dis.shape
    idx = torch.empty((t,) if sync else (t, Np), dtype=torch.int32, device=dis.device)
    mask = torch.empty((t * Np,), dtype=torch.bool, device=dis.device)
    with torch.cuda.device(dis.device):
        N.check(N.lib().rtk_dpselect_select(dis.data_ptr(), T, Np, t, int(sync), idx.data_ptr(), mask.data_ptr(),
                                            N.stream_ptr(dis.device)), "rtk_dpselect_select")
    return idx, mask


def dpselect_gather(x: torch.Tensor, idx: torch.Tensor, sync: bool) -> torch.Tensor:
    """``out[j, p] = x[idx[j, p], p]`` (or ``x[idx[j], p]`` when ``sync``)."""
    N.require_cuda(x, "memory_bank", torch.bfloat16)
    N.require_cuda(idx, "idx", torch.int32)
    x = x.contiguous()
    T, Np, Cc = x.shape
    t = idx.shape[0]
    out = torch.empty((t, Np, Cc), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        N.check(N.lib().rtk_dpselect_gather(x.data_ptr(), T, Np, Cc, idx.data_ptr(), t, int(sync), out.data_ptr(),
                                            N.stream_ptr(x.device)), "rtk_dpselect_gather")
    return out


def memory_bank_compress_keyframe(memory_bank: torch.Tensor, tgt_mem_len: int, window_size: int = 3,
                                  sync: bool = True, return_indices: bool = False) -> tuple:
    """
    Args:
        memory_bank: ``[B=1, T, N, C]`` bf16 CUDA tensor (not modified).
        tgt_mem_len: number of temporal slots to keep (``1 <= t <= T``).
        window_size: arg-rel-max window; the reference only ever passes 3.
        sync: one decision per frame (True) or per patch column (False, the shipped mode).
    Returns:
        compressed_memory_bank ``[1, t, N, C]`` and keypatches_mask ``[t * N]`` (bool); with
        ``return_indices`` additionally the kept indices (int64) the reference computes but drops.
    """
    if memory_bank.dim() != 4:
        raise ValueError("memory_bank must be [B, T, N, C]")
    B, T, Np, Cc = memory_bank.shape
    assert B == 1, "the reference indexes similarity_matrix[0]: batch size 1 only"
    if window_size != 3:
        raise NotImplementedError("window_size != 3 is never used by the reference's callers")
    tgt_mem_len = int(tgt_mem_len)
    if not 1 <= tgt_mem_len <= T:
        raise RuntimeError(f"selected index k out of range (k={tgt_mem_len}, T={T})")     # torch.topk's error
    x = memory_bank[0]
    N.require_cuda(x, "memory_bank", torch.bfloat16)
    x = x.contiguous()
    dev = x.device
    dis = torch.empty((T, Np), dtype=torch.float32, device=dev)
    idx = torch.empty((tgt_mem_len,) if sync else (tgt_mem_len, Np), dtype=torch.int32, device=dev)
    mask = torch.empty((tgt_mem_len * Np,), dtype=torch.bool, device=dev)
    out = torch.empty((1, tgt_mem_len, Np, Cc), dtype=x.dtype, device=dev)
    with torch.cuda.device(dev):                 # distances, selection and compaction: one C-ABI call, three launches
        N.check(N.lib().rtk_dpselect_keyframe(x.data_ptr(), T, Np, Cc, tgt_mem_len, int(bool(sync)), dis.data_ptr(),
                                              idx.data_ptr(), mask.data_ptr(), out.data_ptr(), N.stream_ptr(dev)),
                "rtk_dpselect_keyframe")
    if return_indices:
        return out, mask, idx.long()
    return out, mask


def mallm_compress(memory_bank: torch.Tensor, tgt_mem_len: int, compression_size: torch.Tensor = None, sync: bool = False,
                   hard: bool = False):
    """``memory_bank[1, T, N, C]`` bf16 -> ``([1, t, N, C], sizes [1, t, N] or None)``: the reference's
    ``while bank.shape[1] > tgt_mem_len: bank(, size) = memory_bank_compress_MALLM[_hard](bank(, size), sync)`` loop
    (``qwen2_vl.py:402-409``, ``llava_onevision.py:235-243``) as one fused call, bit-identical for bf16 banks."""
    N.require_cuda(memory_bank, "memory_bank", torch.bfloat16)
    if memory_bank.dim() != 4 or memory_bank.shape[0] != 1:
        raise ValueError("expected memory_bank [1, T, N, C] (the reference only runs batch 1)")
    _, T, Np, Cc = memory_bank.shape
    t = min(int(tgt_mem_len), T)                               # the loop does nothing once T <= tgt
    if t < 1:
        raise ValueError("tgt_mem_len must be >= 1")
    x = memory_bank[0].contiguous()
    sizes_in = None
    if compression_size is not None and not hard:
        N.require_cuda(compression_size, "compression_size", torch.bfloat16)
        if tuple(compression_size.shape) != (1, T, Np):
            raise ValueError("compression_size must be [1, T, N]")
        sizes_in = compression_size[0].contiguous()
    out = torch.empty((1, t, Np, Cc), dtype=x.dtype, device=x.device)
    sizes_out = None if hard else torch.empty((1, t, Np), dtype=x.dtype, device=x.device)
    lib = N.lib()
    ws_bytes = lib.rtk_mallm_workspace_bytes(T, Np, Cc, int(hard))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        N.check(lib.rtk_mallm_compress(x.data_ptr(), sizes_in.data_ptr() if sizes_in is not None else None, T, Np, Cc, t,
                                       int(bool(sync)), int(bool(hard)), out.data_ptr(),
                                       sizes_out.data_ptr() if sizes_out is not None else None, ws.data_ptr(), ws_bytes,
                                       N.stream_ptr(x.device)), "rtk_mallm_compress")
    return out, sizes_out


def memory_bank_compress_MALLM(memory_bank: torch.Tensor, compression_size: torch.Tensor, sync: bool = False):
    """One merge of the most similar adjacent frame pair (``visual_compression.py:5-47``):
    ``([1, T, N, C], [1, T, N]) -> ([1, T-1, N, C], [1, T-1, N])``."""
    if memory_bank.shape[1] < 2:
        raise ValueError("MA-LLM needs at least two frames")
    return mallm_compress(memory_bank, memory_bank.shape[1] - 1, compression_size, sync=sync, hard=False)


def memory_bank_compress_MALLM_hard(memory_bank: torch.Tensor, sync: bool = False):
    """One deletion of the first frame of the most similar adjacent pair (``visual_compression.py:50-83``)."""
    if memory_bank.shape[1] < 2:
        raise ValueError("MA-LLM needs at least two frames")
    return mallm_compress(memory_bank, memory_bank.shape[1] - 1, None, sync=sync, hard=True)[0]
