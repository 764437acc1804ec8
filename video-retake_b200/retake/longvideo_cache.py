"""PivotKV on B200: drop-in for ``retake/longvideo_cache.py`` of SCZwangxiao/video-ReTaKe.

``PivotKVCache`` / ``build_kvcache`` keep the reference's names, constructor, mutable attributes
(``kvcache_compression``, ``keypatches_mask_chunk``, ``pos_embed_reforge``), hooks (``before_forward`` /
``after_forward``), accessors and the ``update(key, value, layer_idx, cache_kwargs)`` contract
(``longvideo_cache.py:119-334``).  A compressing ``update`` is ONE call into the C ABI
(``rtk_pivot_update``, ``include/rtk_b200.h``) that enqueues, on the current stream and without any host
synchronisation, what the reference does with ~40 torch ops and three L x L temporaries:

    rope tables + un-rotation of q, k (reforge)   reference lines 248-259
    tcgen05 Q.K^T + softmax + column sums         lines 260-269
    KV-head mean, key-patch fill, top-k           lines 270-277
    K / V / position gather, temporal re-index    lines 278-295
    re-rotation of the kept keys                  lines 297-306

Storage differs from ``DynamicCache`` (which re-``cat``s the whole past twice per layer and chunk,
``longvideo_cache.py:238,313-318``): every layer owns a geometrically grown buffer; a chunk is appended in
place, ``update`` returns views ``[past | chunk]``, and the kept rows are written over the chunk's head only when
the next cache operation comes in (``update`` of any layer, ``after_forward``, or a read of ``layers[i].keys`` /
``key_cache[i]``) - i.e. after this layer's attention has been enqueued on the stream, because that attention
still needs the uncompressed chunk (``longvideo_cache.py:237,323``).  Single stream, as in the reference.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch
from transformers.cache_utils import Cache, DynamicCache, DynamicLayer
from transformers.utils import logging

from . import _native as N

logger = logging.get_logger(__name__)

__all__ = ["pivot_update", "PivotKVCache", "build_kvcache", "repeat_kv", "rotate_half", "apply_multimodal_rotary_pos_emb",
           "apply_rotary_pos_emb", "pivot_head_scores", "pivot_select", "pivot_compact", "pivot_rope", "pivot_rope_tables"]


# ----------------------------------------------------------------------------------------- thin kernel wrappers
def _hld(x: torch.Tensor, name: str):
    """[1, heads, L, D] (any head/token strides, unit channel stride) -> (tensor, heads, L, D, stride_h, stride_l)."""
    N.require_cuda(x, name, torch.bfloat16)
    if x.dim() != 4 or x.shape[0] != 1:
        raise ValueError(f"{name} must be [1, heads, L, D]")
    if x.stride(3) != 1 or x.data_ptr() % 16 or x.stride(1) % 8 or x.stride(2) % 8:
        x = x.contiguous()
    return x, x.shape[1], x.shape[2], x.shape[3], x.stride(1), x.stride(2)


_WS: Dict[Any, torch.Tensor] = {}


def _workspace(device, need: int) -> torch.Tensor:
    """scratch of the update calls, one per (device, stream): launches on different streams must not share it, and a
    buffer that is replaced by a larger one goes back to the allocator on the stream its kernels were queued on"""
    index = device.index if device.index is not None else torch.cuda.current_device()
    key = (index, torch.cuda.current_stream(index).cuda_stream)
    ws = _WS.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


def _sections(mrope_section):
    return (C.c_int32 * 3)(*[int(s) for s in mrope_section]) if mrope_section else None


def _inv_scale2(attention_scaling: float) -> float:
    # ATen-CUDA divides a bf16 tensor by a python float as x * fp32(1 / fp32(d))
    return float(np.float32(1.0) / np.float32(float(attention_scaling) ** 2))


def pivot_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, mrope_section, attention_scaling: float = 1.0,
               forward: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(Un-)rotate ``x[1, heads, L, D]`` with the rotary tables returned by the model's ``rotary_emb``
    (``[3, 1, L, D]`` with ``mrope_section``, else ``[1, L, D]``; an already selected ``[L, D]`` table also works)."""
    x, heads, L, D, sh, sl = _hld(x, "x")
    cos = cos.contiguous()
    sin = sin.contiguous()
    n_pos = 3 if (mrope_section and cos.numel() == 3 * L * D) else 1
    if cos.numel() != n_pos * L * D or cos.dtype != torch.bfloat16:
        raise ValueError("cos/sin must be bf16 [n_pos, 1, L, D] tables for this chunk")
    if out is None:
        out = torch.empty((1, heads, L, D), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        N.check(N.lib().rtk_pivot_rope(x.data_ptr(), heads, L, D, sh, sl, cos.data_ptr(), sin.data_ptr(), n_pos,
                                       _sections(mrope_section) if n_pos == 3 else None, _inv_scale2(attention_scaling),
                                       int(forward), out.data_ptr(), out.stride(1), out.stride(2),
                                       N.stream_ptr(x.device)), "rtk_pivot_rope")
    return out


def pivot_rope_tables(position_ids: torch.Tensor, inv_freq: torch.Tensor, head_dim: int, mrope_section,
                      attention_scaling: float):
    """bf16 ``[L, D]`` cos/sin tables for ``position_ids`` (``[3, 1, L]`` with mrope, else ``[1, L]``)."""
    N.require_cuda(position_ids, "position_ids", torch.int64)
    L = position_ids.shape[-1]
    pos = position_ids.reshape(-1, L).contiguous()
    n_pos = pos.shape[0]
    inv = inv_freq.to(device=pos.device, dtype=torch.float32).contiguous()
    cos = torch.empty((L, head_dim), dtype=torch.bfloat16, device=pos.device)
    sin = torch.empty_like(cos)
    with torch.cuda.device(pos.device):
        N.check(N.lib().rtk_pivot_rope_tables(pos.data_ptr(), n_pos, L, head_dim, inv.data_ptr(),
                                              _sections(mrope_section) if n_pos == 3 else None, float(attention_scaling),
                                              cos.data_ptr(), sin.data_ptr(), N.stream_ptr(pos.device)),
                "rtk_pivot_rope_tables")
    return cos, sin


def pivot_head_scores(query_states: torch.Tensor, key_states: torch.Tensor) -> torch.Tensor:
    """Per-KV-head pivot scores ``[KVH, L]`` bf16 (``longvideo_cache.py:260-269``)."""
    q, H, L, D, qsh, qsl = _hld(query_states, "query_states")
    k, KVH, Lk, Dk, ksh, ksl = _hld(key_states, "key_states")
    if Lk != L or Dk != D:
        raise ValueError("PivotKV scores the chunk's own keys: key and query lengths must match")
    hs = torch.empty((KVH, L), dtype=torch.bfloat16, device=q.device)
    ws = _workspace(q.device, int(N.lib().rtk_pivot_update_workspace_bytes(H, KVH, L, D)))
    with torch.cuda.device(q.device):
        N.check(N.lib().rtk_pivot_score(q.data_ptr(), H, qsh, qsl, k.data_ptr(), KVH, ksh, ksl, L, D, hs.data_ptr(),
                                        ws.data_ptr(), ws.numel(), N.stream_ptr(q.device)), "rtk_pivot_score")
    return hs


def pivot_select(head_scores: torch.Tensor, keep_len: int, keymask: Optional[torch.Tensor] = None,
                 return_scores: bool = False):
    """Ascending kept indices (int32 ``[keep_len]``) from ``head_scores[KVH, L]`` (``longvideo_cache.py:270-277``)."""
    N.require_cuda(head_scores, "head_scores", torch.bfloat16)
    head_scores = head_scores.contiguous()
    KVH, L = head_scores.shape
    idx = torch.empty((keep_len,), dtype=torch.int32, device=head_scores.device)
    score = torch.empty((L,), dtype=torch.bfloat16, device=head_scores.device) if return_scores else None
    mptr = None
    if keymask is not None:
        N.require_cuda(keymask, "keypatches_mask_chunk", torch.bool)
        keymask = keymask.contiguous()
        if keymask.numel() != L:
            raise ValueError("keypatches_mask_chunk must have one entry per chunk token")
        mptr = keymask.data_ptr()
    with torch.cuda.device(head_scores.device):
        N.check(N.lib().rtk_pivot_select(head_scores.data_ptr(), KVH, L, mptr, keep_len, idx.data_ptr(),
                                         score.data_ptr() if score is not None else None,
                                         N.stream_ptr(head_scores.device)), "rtk_pivot_select")
    return (idx, score) if return_scores else idx


def pivot_compact(key_states: torch.Tensor, value_states: torch.Tensor, keep_idx: torch.Tensor,
                  position_ids: Optional[torch.Tensor] = None, reforge: bool = False):
    """Gather kept K/V rows (and positions) of one chunk (``longvideo_cache.py:278-295``)."""
    k, KVH, L, D, sh, sl = _hld(key_states, "key_states")
    v, _, _, _, vsh, vsl = _hld(value_states, "value_states")
    if (vsh, vsl) != (sh, sl):
        v = v.contiguous()
        k = k.contiguous()
        sh, sl = k.stride(1), k.stride(2)
    keep = keep_idx.numel()
    k_out = torch.empty((1, KVH, keep, D), dtype=k.dtype, device=k.device)
    v_out = torch.empty_like(k_out)
    pos_flat = pos_out = None
    n_pos = 0
    if position_ids is not None:
        pos_flat = position_ids.reshape(-1, L).contiguous()          # [3, L] (mrope) or [1, L]
        n_pos = pos_flat.shape[0]
        pos_out = torch.empty(position_ids.shape[:-1] + (keep,), dtype=torch.int64, device=k.device)
    with torch.cuda.device(k.device):
        N.check(N.lib().rtk_pivot_compact(k.data_ptr(), v.data_ptr(), KVH, L, D, sh, sl, keep_idx.data_ptr(), keep,
                                          k_out.data_ptr(), v_out.data_ptr(), keep * D,
                                          pos_flat.data_ptr() if pos_flat is not None else None, n_pos,
                                          pos_out.data_ptr() if pos_out is not None else None, int(reforge),
                                          N.stream_ptr(k.device)), "rtk_pivot_compact")
    return k_out, v_out, pos_out


def _block_copy(jobs) -> None:
    """``[(src[1, heads, rows, D], dst[1, heads, rows, D]), ...]`` strided bf16 block copies, four per launch"""
    for i in range(0, len(jobs), 4):
        part = jobs[i:i + 4]
        n = len(part)
        srcs, dsts = [], []
        for a, b in part:
            if a.stride(3) != 1 or a.data_ptr() % 16 or a.stride(1) % 8 or a.stride(2) % 8:
                a = a.contiguous()
            srcs.append(a)
            dsts.append(b)
        P, I = C.c_void_p * n, C.c_int64 * n
        dev = dsts[0].device
        with torch.cuda.device(dev):
            N.check(N.lib().rtk_kv_block_copy(
                n, P(*[t.data_ptr() for t in srcs]), P(*[t.data_ptr() for t in dsts]),
                I(*[t.shape[1] for t in srcs]), I(*[t.shape[2] for t in srcs]),
                I(*[t.stride(1) for t in srcs]), I(*[t.stride(2) for t in srcs]),
                I(*[t.stride(1) for t in dsts]), I(*[t.stride(2) for t in dsts]),
                srcs[0].shape[3], N.stream_ptr(dev)), "rtk_kv_block_copy")


class _UpdateArgs(C.Structure):
    """mirror of ``rtk_pivot_update_args`` (include/rtk_b200.h)"""
    _fields_ = [("q", C.c_void_p), ("H", C.c_int64), ("q_stride_h", C.c_int64), ("q_stride_l", C.c_int64),
                ("k", C.c_void_p), ("KVH", C.c_int64), ("k_stride_h", C.c_int64), ("k_stride_l", C.c_int64),
                ("v", C.c_void_p), ("v_stride_h", C.c_int64), ("v_stride_l", C.c_int64),
                ("k_stride_h_in", C.c_int64), ("k_stride_l_in", C.c_int64),
                ("L", C.c_int64), ("D", C.c_int64),
                ("keymask", C.c_void_p), ("keep", C.c_int64),
                ("reforge", C.c_int32), ("n_pos", C.c_int32),
                ("pos", C.c_void_p), ("cos", C.c_void_p), ("sin", C.c_void_p), ("inv_freq", C.c_void_p),
                ("attention_scaling", C.c_float), ("inv_scale2", C.c_float),
                ("mrope_section", C.c_int32 * 3), ("skip_select", C.c_int32),
                ("k_out", C.c_void_p), ("v_out", C.c_void_p), ("out_stride_h", C.c_int64),
                ("pos_out", C.c_void_p), ("keep_idx", C.c_void_p), ("head_scores", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("ev_score_begin", C.c_void_p), ("ev_score_end", C.c_void_p),
                ("pos_out_stride", C.c_int64),
                ("skip_score", C.c_int32), ("score_rows", C.c_int32),
                ("xchg_world", C.c_int32), ("xchg_rank", C.c_int32), ("xchg_epoch", C.c_uint32),
                ("xchg_scores", C.c_void_p * 8), ("xchg_flags", C.c_void_p * 8)]


_STATIC_INV_FREQ_ROPE = ("default", "yarn", "linear", "llama3")


def _rotary_inv_freq(rotary):
    """(inv_freq fp32 tensor, attention_scaling) when ``rotary`` is a stock rotary module with a static inv_freq
    (the tables are then computed inside the library), else None (the module is called like the reference does)."""
    if os.environ.get("RTK_ROTARY_TABLES") == "1":
        return None
    inv = getattr(rotary, "inv_freq", None)
    sc = getattr(rotary, "attention_scaling", None)
    if not isinstance(inv, torch.Tensor) or sc is None or inv.dim() != 1:
        return None
    if getattr(rotary, "rope_type", "default") not in _STATIC_INV_FREQ_ROPE:
        return None
    return inv, float(sc)


# ------------------------------------------------------------------------- helpers kept for API compatibility
def repeat_kv(hidden_states: torch.Tensor, n_rep: int) -> torch.Tensor:
    """(batch, kv_heads, L, D) -> (batch, kv_heads * n_rep, L, D); the kernels index GQA groups instead."""
    b, h, s, d = hidden_states.shape
    if n_rep == 1:
        return hidden_states
    return hidden_states[:, :, None].expand(b, h, n_rep, s, d).reshape(b, h * n_rep, s, d)


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def _rope_pair(q, k, cos, sin, mrope_section, unsqueeze_dim, reverse, attention_scaling):
    if unsqueeze_dim not in (1, 2):
        raise ValueError("unsqueeze_dim must be 1 ([1, heads, L, D]) or 2 ([1, L, heads, D])")
    if reverse and (q is None or k is None):
        raise TypeError("reverse rotation needs both q and k (the reference multiplies both unconditionally)")

    def one(x):
        if x is None:
            return None
        view = x if unsqueeze_dim == 1 else x.transpose(1, 2)
        out = pivot_rope(view, cos, sin, mrope_section, attention_scaling if reverse else 1.0, forward=not reverse)
        return out if unsqueeze_dim == 1 else out.transpose(1, 2)

    return one(q), one(k)


def apply_multimodal_rotary_pos_emb(q, k, cos, sin, mrope_section, unsqueeze_dim=1, reverse=False, attention_scaling=1):
    """Same signature and bf16 rounding chain as ``longvideo_cache.py:36-83`` (``cos`` / ``sin`` ``[3, 1, L, D]`` from the
    model's rotary module; ``reverse=True`` un-rotates and divides by ``attention_scaling ** 2``), on ``rtk_pivot_rope``."""
    return _rope_pair(q, k, cos, sin, list(mrope_section), unsqueeze_dim, reverse, attention_scaling)


def apply_rotary_pos_emb(q, k, cos, sin, position_ids=None, unsqueeze_dim=1, reverse=False, attention_scaling=1):
    """1-D rotary twin of the above (``longvideo_cache.py:86-116``; ``cos`` / ``sin`` ``[1, L, D]``)."""
    return _rope_pair(q, k, cos, sin, None, unsqueeze_dim, reverse, attention_scaling)


def fill_update_args(keymask, reforge, inv_freq_on, query_states, key_states, value_states, position_ids, rotary_emb_fn,
                 mrope_section, keep_len, k_out=None, v_out=None, pos_out=None, own_pos=False):
    """fill one ``rtk_pivot_update_args``; returns (args, outputs, tensors that must outlive the launches, fast)
    ``k_out`` / ``v_out`` ``[1, KVH, keep, D]`` (rows D apart, any head stride) and ``pos_out`` ``[..., keep]`` default to
    fresh tensors; ``fast`` is False when the rotary module is opaque (tables come from calling it)"""
    q, H, L, D, qsh, qsl = _hld(query_states, "query_states")
    k, KVH, Lk, Dk, ksh, ksl = _hld(key_states, "key_states")
    v, _, _, _, vsh, vsl = _hld(value_states, "value_states")
    if Lk != L or Dk != D:
        raise ValueError("PivotKV scores the chunk's own keys: key and query lengths must match")
    dev = q.device
    if k_out is None:
        k_out = torch.empty((1, KVH, keep_len, D), dtype=torch.bfloat16, device=dev)
        v_out = torch.empty_like(k_out)
    head_scores = torch.empty((KVH, L), dtype=torch.bfloat16, device=dev)
    keep_idx = torch.empty((keep_len,), dtype=torch.int32, device=dev)
    a = _UpdateArgs()
    a.q, a.H, a.q_stride_h, a.q_stride_l = q.data_ptr(), H, qsh, qsl
    a.k, a.KVH, a.k_stride_h, a.k_stride_l = k.data_ptr(), KVH, ksh, ksl
    a.v, a.v_stride_h, a.v_stride_l = v.data_ptr(), vsh, vsl
    a.k_stride_h_in, a.k_stride_l_in = ksh, ksl
    a.L, a.D, a.keep = L, D, keep_len
    if keymask is not None:
        N.require_cuda(keymask, "keypatches_mask_chunk", torch.bool)
        keymask = keymask.contiguous()
        if keymask.numel() != L:
            raise ValueError("keypatches_mask_chunk must have one entry per chunk token")
        a.keymask = keymask.data_ptr()
    reforge = bool(reforge)
    a.reforge = int(reforge)
    pos_flat = cos = sin = inv = None
    if position_ids is not None:
        N.require_cuda(position_ids, "position_ids", torch.int64)
        # own_pos (deferred launches): the ids are read at after_forward(), by which time the caller may have
        # re-based the SAME tensor in place for the next layers (the attention forward does) - keep a private copy
        pos_flat = position_ids.reshape(-1, L).clone() if own_pos else position_ids.reshape(-1, L).contiguous()
        a.n_pos, a.pos = pos_flat.shape[0], pos_flat.data_ptr()
        if pos_out is None:
            pos_out = torch.empty(position_ids.shape[:-1] + (keep_len,), dtype=torch.int64, device=dev)
        a.pos_out = pos_out.data_ptr()
        a.pos_out_stride = pos_out.stride(0) if pos_flat.shape[0] > 1 else keep_len
        if mrope_section and pos_flat.shape[0] == 3:
            a.mrope_section = (C.c_int32 * 3)(*[int(s) for s in mrope_section])
    fast = True
    if reforge:
        scaling = float(rotary_emb_fn.attention_scaling)
        a.attention_scaling, a.inv_scale2 = scaling, _inv_scale2(scaling)
        static = _rotary_inv_freq(rotary_emb_fn)
        if static is not None:
            inv = inv_freq_on(static[0], dev)
            a.inv_freq = inv.data_ptr()
        else:
            fast = False
            cos, sin = rotary_emb_fn(value_states, position_ids)
            cos, sin = cos.contiguous(), sin.contiguous()
            if cos.dtype != torch.bfloat16 or cos.numel() != pos_flat.shape[0] * L * D:
                raise ValueError("rotary_emb must return bf16 [n_pos, 1, L, D] tables")
            a.cos, a.sin = cos.data_ptr(), sin.data_ptr()
    a.k_out, a.v_out, a.out_stride_h = k_out.data_ptr(), v_out.data_ptr(), k_out.stride(1)
    a.keep_idx, a.head_scores = keep_idx.data_ptr(), head_scores.data_ptr()
    sig = (H, KVH, L, D, keep_len, reforge, int(a.n_pos), a.inv_freq, float(a.attention_scaling), tuple(a.mrope_section),
           dev.index)
    outs = {"k_out": k_out, "v_out": v_out, "pos_out": pos_out, "head_scores": head_scores, "keep_idx": keep_idx}
    keepalive = (q, k, v, keymask, pos_flat, cos, sin, inv)
    return a, outs, keepalive, fast, sig


_INV_FREQ_DEV = {}


def _inv_freq_on_device(inv_freq: torch.Tensor, dev) -> torch.Tensor:
    """fp32 copy of a rotary module's inv_freq on ``dev`` (cached per tensor)"""
    if inv_freq.device == dev and inv_freq.dtype == torch.float32 and inv_freq.is_contiguous():
        return inv_freq
    key = (id(inv_freq), dev)
    hit = _INV_FREQ_DEV.get(key)
    if hit is None or hit[1] is not inv_freq:
        hit = _INV_FREQ_DEV[key] = (inv_freq.to(device=dev, dtype=torch.float32).contiguous(), inv_freq)
    return hit[0]


def pivot_update(query_states: torch.Tensor, key_states: torch.Tensor, value_states: torch.Tensor, keep_len: int,
                 keymask: Optional[torch.Tensor] = None, position_ids: Optional[torch.Tensor] = None, rotary_emb=None,
                 mrope_section=None, reforge: bool = False, score_events=None):
    """The compressing part of ``PivotKVCache.update`` (``longvideo_cache.py:244-306``) as ONE C-ABI call, without the cache
    bookkeeping: un-rotate, score, select, compact, re-index, re-rotate.
    -> (kept K ``[1, KVH, keep, D]``, kept V, kept positions or None, kept indices int32 ``[keep]``, head scores ``[KVH, L]``)"""
    a, outs, keepalive, fast, _ = fill_update_args(keymask, reforge, _inv_freq_on_device, query_states, key_states,
                                                   value_states, position_ids, rotary_emb, mrope_section, keep_len)
    dev = query_states.device
    lib = N.lib()
    ws = _workspace(dev, int(lib.rtk_pivot_update_workspace_bytes(a.H, a.KVH, a.L, a.D)) + 256)
    ws_ptr = (ws.data_ptr() + 255) & ~255
    a.workspace, a.workspace_bytes = ws_ptr, ws.numel() - (ws_ptr - ws.data_ptr())
    if score_events is not None:
        a.ev_score_begin, a.ev_score_end = score_events
    with torch.cuda.device(dev):
        N.check(lib.rtk_pivot_update(C.byref(a), N.stream_ptr(dev)), "rtk_pivot_update")
    k_out, v_out, pos_out = outs["k_out"], outs["v_out"], outs["pos_out"]
    if reforge and not fast:
        # opaque rotary callable: ask it for the tables of the re-indexed positions, then re-rotate in place
        cos2, sin2 = rotary_emb(v_out, pos_out)
        pivot_rope(k_out, cos2, sin2, mrope_section, 1.0, forward=True, out=k_out)
    del keepalive
    return k_out, v_out, pos_out, outs["keep_idx"], outs["head_scores"]


class PivotKVLayer(DynamicLayer):
    """One layer's K/V in a geometrically grown buffer ``[1, KVH, cap, D]`` with in-place append and a deferred
    overwrite of the last chunk by its kept rows.  ``keys`` / ``values`` are views of the valid prefix."""

    def __init__(self):
        super().__init__()
        self._kbuf = self._vbuf = None
        self._len = 0
        self._pending = None            # (kept_k, kept_v, start) waiting to be written at [start, start + keep)
        self._deferred_owner = None     # PivotKVCache that still owes this layer a deferred (batched) compression

    # -- HF reads these attributes; a read settles the deferred write first
    @property
    def keys(self):
        self.flush()
        return self._keys_view

    @keys.setter
    def keys(self, t):
        self._adopt(t, "k")

    @property
    def values(self):
        self.flush()
        return self._values_view

    @values.setter
    def values(self, t):
        self._adopt(t, "v")

    def _adopt(self, t, which):
        # external assignment (4.48-style ``cache.key_cache[i] = tensor``, crop, reorder ...): take it as the buffer
        if t is None or not isinstance(t, torch.Tensor) or t.numel() == 0:
            if which == "k":
                self._kbuf, self._keys_view, self._len = None, t, 0
            else:
                self._vbuf, self._values_view = None, t
            self._pending = None
            return
        self.flush()
        t = t if t.is_contiguous() else t.contiguous()
        if which == "k":
            self._kbuf, self._keys_view, self._len = t, t, t.shape[-2]
        else:
            self._vbuf, self._values_view = t, t

    def lazy_initialization(self, key_states, value_states=None):
        self.dtype, self.device = key_states.dtype, key_states.device
        self._keys_view = torch.empty((0,), dtype=self.dtype, device=self.device)
        self._values_view = torch.empty((0,), dtype=self.dtype, device=self.device)
        self.is_initialized = True

    def get_seq_length(self) -> int:
        return self._len if self.is_initialized else 0

    def pending_jobs(self):
        """copy jobs that settle the deferred overwrite (and forget it)"""
        p = self._pending
        if p is None:
            return []
        self._pending = None
        kk, vv, start = p
        n = kk.shape[2]
        return [(kk, self._kbuf[:, :, start:start + n]), (vv, self._vbuf[:, :, start:start + n])]

    def flush(self):
        owner = self._deferred_owner
        if owner is not None:
            owner.flush_deferred()      # the batched compression of the last chunk writes this layer's kept rows
        jobs = self.pending_jobs()
        if jobs:
            _block_copy(jobs)

    def _reserve(self, heads, n, d):
        cap = 0 if self._kbuf is None else self._kbuf.shape[2]
        if n <= cap and self._vbuf is not None and self._vbuf.shape[2] >= n:
            return
        new_cap = max(n, int(cap * 1.5), 8192)
        kb = torch.empty((1, heads, new_cap, d), dtype=self.dtype, device=self.device)
        vb = torch.empty_like(kb)
        if self._len:
            kb[:, :, :self._len].copy_(self._kbuf[:, :, :self._len])
            vb[:, :, :self._len].copy_(self._vbuf[:, :, :self._len])
        self._kbuf, self._vbuf = kb, vb

    def append_jobs(self, key_states, value_states):
        """reserve room, advance the length and return (copy jobs, keys view, values view) for ``[past | new]``"""
        if not self.is_initialized:
            self.lazy_initialization(key_states, value_states)
        _, heads, n_new, d = key_states.shape
        p = self._len
        cap = 0 if self._kbuf is None else self._kbuf.shape[2]
        if p + n_new > cap:
            self.flush()                                  # the old buffer is about to be copied: settle it first
            self._reserve(heads, p + n_new, d)
        jobs = [(key_states, self._kbuf[:, :, p:p + n_new]), (value_states, self._vbuf[:, :, p:p + n_new])]
        self._len = p + n_new
        self._keys_view = self._kbuf[:, :, :self._len]
        self._values_view = self._vbuf[:, :, :self._len]
        return jobs, self._keys_view, self._values_view

    def update(self, key_states, value_states, *args, **kwargs):
        """append in place, return views ``[past | new]``"""
        # room first: if the buffer has to grow, append_jobs settles the pending overwrite into the OLD buffer before it
        # is copied; popping the pending jobs earlier would leave them pointing at the discarded buffer
        more, k_all, v_all = self.append_jobs(key_states, value_states)
        jobs = self.pending_jobs()
        _block_copy(jobs + more)
        return k_all, v_all

    def replace_tail(self, n_tail, kept_k, kept_v):
        """the last ``n_tail`` rows become ``kept_*`` - length changes now, bytes move at the next flush"""
        start = self._len - n_tail
        self._pending = (kept_k, kept_v, start)
        self._len = start + kept_k.shape[2]
        self._keys_view = self._kbuf[:, :, :self._len]
        self._values_view = self._vbuf[:, :, :self._len]

    def shrink_tail(self, n_tail, keep, owner):
        """deferred compression: the last ``n_tail`` rows will become ``keep`` rows written IN PLACE by ``owner``'s batched
        compression.  Returns the (keys, values) destination views ``[1, KVH, keep, D]`` inside the buffers."""
        start = self._len - n_tail
        self._len = start + keep
        self._keys_view = self._kbuf[:, :, :self._len]
        self._values_view = self._vbuf[:, :, :self._len]
        self._deferred_owner = owner
        return self._kbuf[:, :, start:start + keep], self._vbuf[:, :, start:start + keep]


class _LayerListView:
    """``cache.key_cache[i]`` / ``cache.value_cache[i]`` of transformers 4.48 on top of ``layers[i]``."""

    def __init__(self, cache, attr):
        self._c, self._a = cache, attr

    def __getitem__(self, i):
        return getattr(self._c.layers[i], self._a)

    def __setitem__(self, i, v):
        setattr(self._c.layers[i], self._a, v)

    def __len__(self):
        return len(self._c.layers)

    def __iter__(self):
        return (getattr(l, self._a) for l in self._c.layers)


class PivotKVCache(DynamicCache):
    def __init__(self, config) -> None:
        Cache.__init__(self, layer_class_to_replicate=PivotKVLayer)
        self.config = config
        llm_config = config.text_config if hasattr(config, "text_config") else config   # LLaVA-OneVision / Qwen2-VL
        self.hidden_size = llm_config.hidden_size
        self.num_hidden_layers = llm_config.num_hidden_layers
        self.num_heads = llm_config.num_attention_heads
        self.head_dim = self.hidden_size // self.num_heads
        self.num_key_value_heads = llm_config.num_key_value_heads
        self.num_key_value_groups = self.num_heads // self.num_key_value_heads

        kv_compression_kwargs = config.longvideo_kwargs["kvcache_compression_kwargs"]
        self.kvcache_compression = True
        self.kv_compression_kwargs = kv_compression_kwargs
        self.compression_ratio = kv_compression_kwargs["compression_ratio"]
        self.compression_method = kv_compression_kwargs["compression_method"]
        self.pos_embed_reforge = kv_compression_kwargs.get("pos_embed_reforge", False)
        self.position_cache: List[torch.Tensor] = []
        self.num_evicted_tokens: List[int] = []
        self.keypatches_mask_chunk: Optional[torch.Tensor] = None
        self._pos_buf: List[Optional[torch.Tensor]] = []
        self._pos_len: List[int] = []
        self._dirty: List[PivotKVLayer] = []
        self.score_events = None            # optional (cudaEvent_t, cudaEvent_t) ints recorded around the next scoring
        # Deferred compression: ``update`` only appends the chunk and remembers its tensors; the
        # compression of ALL layers of the chunk runs as one batched call (``rtk_pivot_update_batch``, seven launches
        # per chunk - nine with a key-patch mask - instead of eight per layer) when the chunk's forward is over - ``after_forward()``, the hook the
        # reference's chunk loop already calls (``qwen2_vl.py:715-716``) - or at the next cache operation that needs
        # the result.  A chunk's kept rows are first read by the NEXT chunk, so the cache contents every reader sees
        # are the same as with compression inside ``update``.  Knob: ``kvcache_compression_kwargs.deferred_compression``
        # (default: the RTK_DEFERRED environment variable, off).
        self.deferred_compression = bool(kv_compression_kwargs.get("deferred_compression",
                                                                   os.environ.get("RTK_DEFERRED", "0") == "1"))
        self._deferred: List[Dict[str, Any]] = []
        # exposed for tests / the multi-GPU path: the last chunk's per-KV-head scores and kept indices
        self._last_head_scores: Optional[torch.Tensor] = None
        self._last_keep_indices: Optional[torch.Tensor] = None

    @property
    def last_head_scores(self):
        self.flush_deferred()
        return self._last_head_scores

    @last_head_scores.setter
    def last_head_scores(self, t):
        self._last_head_scores = t

    @property
    def last_keep_indices(self):
        self.flush_deferred()
        return self._last_keep_indices

    @last_keep_indices.setter
    def last_keep_indices(self, t):
        self._last_keep_indices = t

    # transformers 4.48 attribute names
    @property
    def key_cache(self):
        return _LayerListView(self, "keys")

    @property
    def value_cache(self):
        return _LayerListView(self, "values")

    def before_forward(self, **kwargs):
        pass

    def after_forward(self, **kwargs):
        self.flush()

    def flush(self):
        """settle every deferred compression / tail overwrite (cheap no-op when nothing is pending)"""
        self.flush_deferred()
        dirty, self._dirty = self._dirty, []
        for layer in dirty:
            layer.flush()

    def flush_deferred(self):
        """run the batched compression of every layer that ``update`` left pending (one C-ABI call per group of layers
        with the same shape - in practice one per chunk)"""
        pend, self._deferred = self._deferred, []
        if not pend:
            return
        lib = N.lib()
        groups: Dict[Any, List[Dict[str, Any]]] = {}
        for e in pend:
            groups.setdefault(e["sig"], []).append(e)
        try:
            for sig, entries in groups.items():
                H, KVH, L, D = sig[:4]
                dev = entries[0]["device"]
                n = len(entries)
                arr = (_UpdateArgs * n)(*[e["args"] for e in entries])
                ws = _workspace(dev, int(lib.rtk_pivot_update_batch_workspace_bytes(H, KVH, L, D, n)) + 256)
                ws_ptr = (ws.data_ptr() + 255) & ~255
                if self.score_events is not None:
                    arr[0].ev_score_begin, arr[0].ev_score_end = self.score_events
                    self.score_events = None
                with torch.cuda.device(dev):
                    N.check(lib.rtk_pivot_update_batch(arr, n, ws_ptr, ws.numel() - (ws_ptr - ws.data_ptr()),
                                                       N.stream_ptr(dev)), "rtk_pivot_update_batch")
        finally:
            # whatever happened, no layer may keep pointing at a batch that will never run again (a failed launch
            # raises to the caller; a later read of the cache must not recurse into this queue)
            for e in pend:
                e["layer"]._deferred_owner = None

    def update_num_evicted_tokens(self, num_tokens: int, layer_idx: int) -> int:
        while len(self.num_evicted_tokens) <= layer_idx:
            self.num_evicted_tokens.append(0)
        self.num_evicted_tokens[layer_idx] += num_tokens
        return self.num_evicted_tokens[layer_idx]

    def _position_slots(self, like: torch.Tensor, n: int, layer_idx: int) -> torch.Tensor:
        """``n`` more entries along the last dim of layer ``layer_idx``'s grown position buffer (leading dims, dtype and
        device of ``like``); returns the view to fill, ``position_cache[layer_idx]`` becomes the longer valid view"""
        while len(self.position_cache) <= layer_idx:
            self.position_cache.append([])
            self._pos_buf.append(None)
            self._pos_len.append(0)
        buf, cur = self._pos_buf[layer_idx], self._pos_len[layer_idx]
        if buf is None or cur + n > buf.shape[-1] or buf.shape[:-1] != like.shape[:-1]:
            if layer_idx < len(self.layers) and self.layers[layer_idx]._deferred_owner is not None:
                self.flush_deferred()                      # this layer's pending batched write targets the old buffer
            cap = max(cur + n, int((0 if buf is None else buf.shape[-1]) * 1.5), 8192)
            nb = torch.empty(like.shape[:-1] + (cap,), dtype=like.dtype, device=like.device)
            if cur:
                nb[..., :cur].copy_(buf[..., :cur])
            buf = self._pos_buf[layer_idx] = nb
        self._pos_len[layer_idx] = cur + n
        self.position_cache[layer_idx] = buf[..., :cur + n]
        return buf[..., cur:cur + n]

    def update_position_ids(self, position_ids: torch.Tensor, layer_idx: int) -> torch.Tensor:
        """append along the last dim into a grown buffer; ``position_cache[layer_idx]`` is the valid view"""
        self._position_slots(position_ids, position_ids.shape[-1], layer_idx).copy_(position_ids)
        return self.position_cache[layer_idx]

    def get_prev_temporal_idx(self, layer_idx: int):
        if len(self.position_cache) <= layer_idx or len(self.position_cache[layer_idx]) == 0:
            return -1
        if layer_idx < len(self.layers) and self.layers[layer_idx]._deferred_owner is not None:
            self.flush_deferred()                       # (the attention forward asks per layer: only this layer's debt matters)
        cache_layer = self.position_cache[layer_idx]
        return cache_layer[0, 0, -1] if cache_layer.ndim == 3 else cache_layer[0, -1]

    def rebased_position_ids(self, position_ids: torch.Tensor, n_layers: int) -> torch.Tensor:
        """The ids a new chunk gets in EVERY layer, in a handful of launches: the attention forward continues the temporal
        row after each layer's compacted cache (``ids[0] += prev_l + 1 - ids[0, ..., 0]``, reference ``qwen2_vl.py:68-73`` /
        ``llava_onevision.py:80-89``); at the start of a chunk the previous chunk is settled in all layers, so the ``prev_l``
        are all known.  ``position_ids`` ``[3, 1, L]`` or ``[1, L]`` -> ``[n_layers, 3, 1, L]`` / ``[n_layers, 1, L]``; layer
        l's slice is a tensor of its own (it can be handed to ``update`` with ``position_ids_owned``)."""
        prevs = [self.get_prev_temporal_idx(l) for l in range(n_layers)]
        dev = position_ids.device
        # host integers (-1: nothing cached yet) become device values through fill kernels: torch.tensor(list, device=cuda)
        # is a pageable host-to-device copy, which synchronises the stream - once per video, with the GPU idle behind it
        if all(not isinstance(p, torch.Tensor) for p in prevs) and len(set(int(p) for p in prevs)) == 1:
            prev = torch.full((n_layers,), int(prevs[0]), dtype=position_ids.dtype, device=dev)
        else:
            prev = torch.stack([p if isinstance(p, torch.Tensor) else torch.full((), int(p), dtype=position_ids.dtype, device=dev)
                                for p in prevs])
        out = position_ids.unsqueeze(0).repeat(n_layers, *([1] * position_ids.dim()))
        first = position_ids[0].reshape(-1)[0]
        t = out[:, 0]                                                  # the temporal rows of all layers
        t += (prev + (1 - first)).view(-1, *([1] * (t.dim() - 1)))
        return out

    def get_seq_length(self, layer_idx: int = 0) -> int:
        if layer_idx >= len(self.layers):
            return 0
        return self.layers[layer_idx].get_seq_length()

    # ------------------------------------------------------------------------------------------------ update
    def _update_args(self, query_states, key_states, value_states, position_ids, rotary_emb_fn, mrope_section, keep_len,
                     k_out=None, v_out=None, pos_out=None, own_pos=False):
        return fill_update_args(getattr(self, "keypatches_mask_chunk", None), self.pos_embed_reforge, self._inv_freq_on,
                                query_states, key_states, value_states, position_ids, rotary_emb_fn, mrope_section, keep_len,
                                k_out, v_out, pos_out, own_pos)

    def _inv_freq_on(self, inv_freq: torch.Tensor, dev) -> torch.Tensor:
        """fp32 copy of the rotary module's inv_freq on ``dev`` (the module's own tensor when it already is one)"""
        return _inv_freq_on_device(inv_freq, dev)

    def _compress_chunk(self, query_states, key_states, value_states, position_ids, rotary_emb_fn, mrope_section,
                        keep_len):
        """one ``rtk_pivot_update`` call -> (kept K [1,KVH,keep,D], kept V, kept positions or None)"""
        ev, self.score_events = self.score_events, None
        k_out, v_out, pos_out, keep_idx, head_scores = pivot_update(
            query_states, key_states, value_states, keep_len, getattr(self, "keypatches_mask_chunk", None), position_ids,
            rotary_emb_fn, mrope_section, self.pos_embed_reforge, ev)
        self.last_head_scores, self.last_keep_indices = head_scores, keep_idx
        return k_out, v_out, pos_out

    def _defer_chunk(self, layer, layer_idx, query_states, key_states, value_states, position_ids, rotary_emb_fn,
                     mrope_section, keep_len, own_pos=True) -> bool:
        """queue this layer's compression for the batched call; False when it has to run now (opaque rotary module)"""
        if self.pos_embed_reforge and _rotary_inv_freq(rotary_emb_fn) is None:
            return False
        # the batched kernels' envelope (rtk_pivot_update_batch): outside it the chunk takes the immediate path, whose
        # C-ABI call reports the problem BEFORE any cache state has changed
        H, KVH, D = query_states.shape[1], key_states.shape[1], query_states.shape[3]
        if (query_states.shape[2] > 16384 or D not in (64, 128) or H % KVH != 0 or query_states.dtype != torch.bfloat16
                or key_states.shape[2] != query_states.shape[2]):
            return False
        q_len = query_states.shape[2]
        pos_out = None
        if self.pos_embed_reforge:
            pos_out = self._position_slots(position_ids, keep_len, layer_idx)      # kept positions land in the position cache
        k_dst, v_dst = layer.shrink_tail(q_len, keep_len, self)                    # kept rows land in the cache itself
        a, outs, keepalive, _, sig = self._update_args(query_states, key_states, value_states, position_ids, rotary_emb_fn,
                                                       mrope_section, keep_len, k_dst, v_dst, pos_out, own_pos=own_pos)
        self._deferred.append({"args": a, "outs": outs, "keepalive": keepalive, "sig": sig, "layer": layer,
                               "device": query_states.device})
        self._last_head_scores, self._last_keep_indices = outs["head_scores"], outs["keep_idx"]
        return True

    def update(self, key_states: torch.Tensor, value_states: torch.Tensor, layer_idx: int,
               cache_kwargs: Optional[Dict[str, Any]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """
        key_states/value_states ``[1, KVH, L, D]``; ``cache_kwargs`` carries ``query_states [1, H, L, D]``,
        ``position_ids`` (``[3, 1, L]`` / ``[1, L]``), ``rotary_emb`` and ``mrope_section`` (consumed).
        Returns the UNcompressed ``[past | chunk]`` keys/values for this step's attention; the cache itself
        keeps ``[past | kept]``.
        """
        logger.warning_once("Enable PivotKVCache compression: length after compression %.2f" % (self.compression_ratio))
        cache_kwargs = cache_kwargs if cache_kwargs is not None else {}
        position_ids = cache_kwargs.pop("position_ids", None)
        # optional (not in the reference): the caller hands this tensor over - it will not be modified before after_forward() -
        # so that deferred compression need not take a private copy of the ids
        pos_owned = bool(cache_kwargs.pop("position_ids_owned", False))

        # 1) this chunk attends to everything: [past | chunk] is what the caller gets back.  The previous layer's
        #    attention is on the stream by now, so its kept rows may land: both go out as ONE block-copy launch.
        while len(self.layers) <= layer_idx:
            self.layers.append(PivotKVLayer())
        layer = self.layers[layer_idx]
        if layer._deferred_owner is not None:
            self.flush_deferred()                       # no after_forward() since this layer's last chunk: settle it now
        # (room first - see PivotKVLayer.update: a growing buffer settles this layer's own pending overwrite itself)
        more, key_states_output, value_states_output = layer.append_jobs(key_states, value_states)
        dirty, self._dirty = self._dirty, []
        jobs = [j for l in dirty for j in l.pending_jobs()]
        _block_copy(jobs + more)

        if self.kvcache_compression:
            query_states = cache_kwargs.pop("query_states")
            rotary_emb_fn = cache_kwargs.pop("rotary_emb", None)
            mrope_section = cache_kwargs.pop("mrope_section", None)
            bsz, num_heads, q_len, head_dim = query_states.shape
            num_key_value_heads, k_len = key_states.shape[1:3]
            assert bsz == 1
            if self.pos_embed_reforge and (position_ids is None or rotary_emb_fn is None):
                raise ValueError("pos_embed_reforge needs position_ids and rotary_emb in cache_kwargs")

            # 2) score the chunk's own keys with the chunk's queries, keep the top ratio * q_len
            keep_len = max(1, int(self.compression_ratio * q_len))
            if self.deferred_compression and self._defer_chunk(layer, layer_idx, query_states, key_states, value_states,
                                                               position_ids, rotary_emb_fn, mrope_section, keep_len,
                                                               own_pos=not pos_owned):
                # 2') ... later: all layers of the chunk in one batched call, kept rows written in place (flush_deferred)
                self.update_num_evicted_tokens(k_len - keep_len, layer_idx)
                return key_states_output, value_states_output
            kept_k, kept_v, kept_pos = self._compress_chunk(query_states, key_states, value_states, position_ids,
                                                            rotary_emb_fn, mrope_section, keep_len)
            if self.pos_embed_reforge:
                self.update_position_ids(kept_pos, layer_idx)
            self.update_num_evicted_tokens(k_len - keep_len, layer_idx)

            # 3) cache keeps [past | kept]
            layer.replace_tail(q_len, kept_k, kept_v)
            self._dirty.append(layer)
        else:
            if self.pos_embed_reforge:
                self.update_position_ids(position_ids, layer_idx)

        return key_states_output, value_states_output


def build_kvcache(config):
    if getattr(config, "longvideo_kwargs", None) is None or not config.longvideo_kwargs.get("kvcache_compression", False):
        return DynamicCache()
    compression_method = config.longvideo_kwargs["kvcache_compression_kwargs"]["compression_method"]
    if compression_method.lower() == "pivotkv":
        return PivotKVCache(config)
    raise NotImplementedError
