"""PivotKV on B200: drop-in for ``retake/longvideo_cache.py`` of SCZwangxiao/video-ReTaKe.

``PivotKVCache`` / ``build_kvcache`` keep the reference's names, constructor, mutable attributes
(``kvcache_compression``, ``keypatches_mask_chunk``, ``pos_embed_reforge``), hooks (``before_forward`` /
``after_forward``), accessors and the ``update(key, value, layer_idx, cache_kwargs)`` contract
(``longvideo_cache.py:119-334``).  What ``update`` does when compression is on is five kernel launches
through the C ABI (``include/rtk_b200.h``) on the current stream instead of ~40 torch ops and three L x L
temporaries:

    rtk_pivot_rope (x2, reforge only)  un-rotate q and k                       (reference lines 248-259)
    rtk_pivot_score                    tcgen05 Q.K^T + softmax + column sums   (lines 260-269)
    rtk_pivot_select                   KV-head mean, key-patch fill, top-k     (lines 270-277)
    rtk_pivot_compact                  K/V/position gather, temporal re-index  (lines 278-295)
    rtk_pivot_rope (forward, reforge)  re-rotate the kept keys                 (lines 297-306)

The cache object is a ``transformers.DynamicCache`` (5.x layout: ``layers[i].keys/.values``);
``key_cache`` / ``value_cache`` are kept as list views for code written against 4.48.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch
from transformers.cache_utils import DynamicCache
from transformers.utils import logging

from . import _native as N

logger = logging.get_logger(__name__)

__all__ = ["PivotKVCache", "build_kvcache", "repeat_kv", "rotate_half", "pivot_head_scores", "pivot_select",
           "pivot_compact", "pivot_rope"]


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


def _workspace(device, H: int, L: int) -> torch.Tensor:
    need = int(N.lib().rtk_pivot_score_workspace_bytes(H, L))
    key = (device.index if device.index is not None else torch.cuda.current_device())
    ws = _WS.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


def pivot_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, mrope_section, attention_scaling: float = 1.0,
               forward: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(Un-)rotate ``x[1, heads, L, D]`` with the rotary tables returned by the model's ``rotary_emb``."""
    x, heads, L, D, sh, sl = _hld(x, "x")
    n_pos = 3 if mrope_section else 1
    cos = cos.contiguous()
    sin = sin.contiguous()
    if cos.numel() != n_pos * L * D or cos.dtype != torch.bfloat16:
        raise ValueError("cos/sin must be bf16 [n_pos, 1, L, D] tables for this chunk")
    if out is None:
        out = torch.empty((1, heads, L, D), dtype=x.dtype, device=x.device)
    sec = (C.c_int32 * 3)(*[int(s) for s in mrope_section]) if mrope_section else None
    s2 = np.float32(float(attention_scaling) ** 2)
    inv = float(np.float32(1.0) / s2)
    with torch.cuda.device(x.device):
        N.check(N.lib().rtk_pivot_rope(x.data_ptr(), heads, L, D, sh, sl, cos.data_ptr(), sin.data_ptr(), n_pos, sec,
                                       inv, int(forward), out.data_ptr(), out.stride(1), out.stride(2),
                                       N.stream_ptr(x.device)), "rtk_pivot_rope")
    return out


def pivot_head_scores(query_states: torch.Tensor, key_states: torch.Tensor) -> torch.Tensor:
    """Per-KV-head pivot scores ``[KVH, L]`` bf16 (``longvideo_cache.py:260-269``)."""
    q, H, L, D, qsh, qsl = _hld(query_states, "query_states")
    k, KVH, Lk, Dk, ksh, ksl = _hld(key_states, "key_states")
    if Lk != L or Dk != D:
        raise ValueError("PivotKV scores the chunk's own keys: key and query lengths must match")
    hs = torch.empty((KVH, L), dtype=torch.bfloat16, device=q.device)
    ws = _workspace(q.device, H, L)
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
        super().__init__()
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
        # exposed for tests / the multi-GPU path: last chunk's per-KV-head scores and kept indices
        self.last_head_scores: Optional[torch.Tensor] = None
        self.last_keep_indices: Optional[torch.Tensor] = None

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
        pass

    def update_num_evicted_tokens(self, num_tokens: int, layer_idx: int) -> int:
        while len(self.num_evicted_tokens) <= layer_idx:
            self.num_evicted_tokens.append(0)
        self.num_evicted_tokens[layer_idx] += num_tokens
        return self.num_evicted_tokens[layer_idx]

    def update_position_ids(self, position_ids: torch.Tensor, layer_idx: int) -> torch.Tensor:
        while len(self.position_cache) < layer_idx:
            self.position_cache.append([])
        if len(self.position_cache) == layer_idx:
            self.position_cache.append(position_ids)
        elif len(self.position_cache[layer_idx]) == 0:
            self.position_cache[layer_idx] = position_ids
        else:
            self.position_cache[layer_idx] = torch.cat([self.position_cache[layer_idx], position_ids], dim=-1)
        return self.position_cache[layer_idx]

    def get_prev_temporal_idx(self, layer_idx: int):
        if len(self.position_cache) <= layer_idx:
            return -1
        cache_layer = self.position_cache[layer_idx]
        return cache_layer[0, 0, -1] if cache_layer.ndim == 3 else cache_layer[0, -1]

    def select_keep_indices(self, head_scores: torch.Tensor, keep_len: int) -> torch.Tensor:
        """Hook between scoring and selection; the KV-head-sharded cache all-gathers ``head_scores`` here."""
        return pivot_select(head_scores, keep_len, getattr(self, "keypatches_mask_chunk", None))

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

        # 1) this chunk attends to everything: [past | chunk] is what the caller gets back
        key_states_output, value_states_output = super().update(key_states, value_states, layer_idx)

        if self.kvcache_compression:
            query_states = cache_kwargs.pop("query_states")
            rotary_emb_fn = cache_kwargs.pop("rotary_emb", None)
            mrope_section = cache_kwargs.pop("mrope_section", None)
            bsz, num_heads, q_len, head_dim = query_states.shape
            num_key_value_heads, k_len = key_states.shape[1:3]
            assert bsz == 1
            if self.pos_embed_reforge and position_ids is None:
                raise ValueError("pos_embed_reforge needs position_ids in cache_kwargs")

            if self.pos_embed_reforge:
                cos, sin = rotary_emb_fn(value_states, position_ids)
                scaling = rotary_emb_fn.attention_scaling
                query_states = pivot_rope(query_states, cos, sin, mrope_section, scaling, forward=False)
                key_states = pivot_rope(key_states, cos, sin, mrope_section, scaling, forward=False)

            # 2) score the chunk's own keys with the chunk's queries, keep the top ratio * q_len
            keep_len = max(1, int(self.compression_ratio * q_len))
            head_scores = pivot_head_scores(query_states, key_states)
            keep_indices = self.select_keep_indices(head_scores, keep_len)
            self.last_head_scores, self.last_keep_indices = head_scores, keep_indices

            compressed_key_states, compressed_value_states, compressed_position_ids = pivot_compact(
                key_states, value_states, keep_indices, position_ids, reforge=self.pos_embed_reforge)

            if self.pos_embed_reforge:
                cos, sin = rotary_emb_fn(compressed_value_states, compressed_position_ids)
                pivot_rope(compressed_key_states, cos, sin, mrope_section, 1.0, forward=True, out=compressed_key_states)
                self.update_position_ids(compressed_position_ids, layer_idx)
            self.update_num_evicted_tokens(k_len - keep_len, layer_idx)

            # 3) cache keeps [past | kept]
            layer = self.layers[layer_idx]
            layer.keys = torch.cat([key_states_output[..., :-q_len, :], compressed_key_states], dim=2)
            layer.values = torch.cat([value_states_output[..., :-q_len, :], compressed_value_states], dim=2)
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
