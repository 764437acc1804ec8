"""Oracle for PivotKV (chunk-local pivot scoring, top-k, KV compaction).  TEST INFRASTRUCTURE ONLY.

Restates ``retake/longvideo_cache.py:16-334`` with every bf16 rounding written out
(notes N3/N4).  ``semantics`` selects which ATen device rules are replayed
where CPU and CUDA differ:

=========================  ================================  ===============================
step (longvideo_cache.py)  ``semantics="cuda"``              ``semantics="cpu"``
=========================  ================================  ===============================
``/ sqrt(D)``      (:264)  ``bf16(s * f32(1/f32(sqrt D)))``  ``bf16(s / f32(sqrt D))``
``.mean(1)``       (:269)  ``bf16(sum_f32 * f32(1/G))``      ``bf16(sum_f32 / G)``
``.mean(0)``       (:270)  ``bf16(sum_f32 * f32(1/KVH))``    ``bf16(sum_f32 / KVH)``
``/ scaling**2``   (:76)   ``bf16(x * f32(1/f32(s*s)))``     ``bf16(x / f32(s*s))``
=========================  ================================  ===============================

The Q.K^T contraction itself is ``torch.matmul`` in fp32 on the bf16 values and then
rounded to bf16 - the fp32 accumulation order of cuBLAS / oneDNN is not reproducible,
which is why scores carry a tolerance (1e-2 relative, BASELINE.json) and kept indices are
compared tie/ulp-aware whenever the scores were not produced by the same matmul.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch

BF16 = torch.bfloat16


def _r(x: torch.Tensor) -> torch.Tensor:
    return x.to(BF16).to(torch.float32)


def _rounder(dtype):
    return _r if dtype == BF16 else (lambda t: t)


def rotate_half_f32(x: torch.Tensor) -> torch.Tensor:
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def select_mrope(cos: torch.Tensor, mrope_section: Optional[List[int]]) -> torch.Tensor:
    """[3, 1, L, D] -> [1, L, D] picking temporal/height/width per channel block
    (``longvideo_cache.py:67-73``); a [1, L, D] table passes through."""
    if not mrope_section:
        return cos
    sec = list(mrope_section) * 2
    parts = cos.split(sec, dim=-1)
    return torch.cat([m[i % 3] for i, m in enumerate(parts)], dim=-1)


def unrotate(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, scaling: float,
             semantics: str = "cuda") -> torch.Tensor:
    """``reverse=True`` branch (``longvideo_cache.py:75-77,108-110``) on x[1, h, L, D];
    cos/sin already [1, L, D] in x.dtype."""
    rnd = _rounder(x.dtype)
    xf, c, s = x.float(), cos.float()[:, None], sin.float()[:, None]
    a = rnd(xf * c)
    b = rnd(rnd(rotate_half_f32(xf)) * s)
    d = rnd(a - b)
    s2 = torch.tensor(scaling ** 2, dtype=torch.float32)
    if semantics == "cuda":
        out = rnd(d * (torch.tensor(1.0, dtype=torch.float32) / s2))
    else:
        out = rnd(d / s2)
    return out.to(x.dtype)


def rotate(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """forward branch (``longvideo_cache.py:79-81,112-114``)."""
    rnd = _rounder(x.dtype)
    xf, c, s = x.float(), cos.float()[:, None], sin.float()[:, None]
    out = rnd(rnd(xf * c) + rnd(rnd(rotate_half_f32(xf)) * s))
    return out.to(x.dtype)


def pivot_scores(q: torch.Tensor, k: torch.Tensor, semantics: str = "cuda",
                 return_partials: bool = False):
    """score[L] in q.dtype from q[1, H, L, D], k[1, KVH, L, D] (``longvideo_cache.py:260-270``)."""
    _, H, L, D = q.shape
    KVH = k.shape[1]
    G = H // KVH
    rnd = _rounder(q.dtype)
    one = torch.tensor(1.0, dtype=torch.float32)
    kr = k[:, :, None].expand(1, KVH, G, L, D).reshape(1, H, L, D)
    s = rnd(torch.matmul(q.float(), kr.float().transpose(2, 3)))
    sq = torch.tensor(math.sqrt(D), dtype=torch.float32)
    s = rnd(s * (one / sq)) if semantics == "cuda" else rnd(s / sq)
    m = s.max(-1, keepdim=True).values
    e = torch.exp(s - m)
    p = rnd(e / e.sum(-1, keepdim=True))
    a = rnd(p[0].sum(1))                                    # [H, L]   sum over queries
    gsum = a.reshape(KVH, G, L).sum(1)
    b = rnd(gsum * (one / G)) if semantics == "cuda" else rnd(gsum / G)       # [KVH, L]
    ksum = b.sum(0)
    score = rnd(ksum * (one / KVH)) if semantics == "cuda" else rnd(ksum / KVH)  # [L]
    if return_partials:
        return score.to(q.dtype), a.to(q.dtype), b.to(q.dtype)
    return score.to(q.dtype)


def keep_indices(score: torch.Tensor, keep_len: int, keymask: Optional[torch.Tensor] = None,
                 tie: str = "lowest") -> torch.Tensor:
    """(``longvideo_cache.py:272-277``) mask-fill with 1.0, top-k, ascending sort."""
    s = score.clone()
    if keymask is not None:
        s = s.masked_fill(keymask, 1.0)
    if tie == "torch":
        idx = s.topk(keep_len).indices
    else:
        idx = torch.sort(s.float(), descending=True, stable=True).indices[:keep_len]
    return idx.sort().values


def reforge_temporal(pos_t: torch.Tensor, keep_len: int, k_len: int) -> torch.Tensor:
    """(``longvideo_cache.py:293-295``)  m + ((p - m) * (keep/k_len)).long(); the product is
    int64 * python float -> float32 tensor arithmetic, truncated toward zero."""
    m = pos_t.min()
    ratio = keep_len / k_len
    return m + ((pos_t - m) * ratio).long()


class OraclePivotKVCache:
    """Plain-Python replay of ``PivotKVCache`` (``longvideo_cache.py:119-323``) without the HF base
    class: per-layer ``key_cache`` / ``value_cache`` / ``position_cache`` / ``num_evicted_tokens``."""

    def __init__(self, num_heads: int, num_kv_heads: int, head_dim: int, compression_ratio: float,
                 pos_embed_reforge: bool = False, semantics: str = "cuda", tie: str = "lowest"):
        self.num_heads, self.num_key_value_heads, self.head_dim = num_heads, num_kv_heads, head_dim
        self.compression_ratio = compression_ratio
        self.pos_embed_reforge = pos_embed_reforge
        self.kvcache_compression = True
        self.keypatches_mask_chunk = None
        self.semantics, self.tie = semantics, tie
        self.key_cache: List[torch.Tensor] = []
        self.value_cache: List[torch.Tensor] = []
        self.position_cache: List[torch.Tensor] = []
        self.num_evicted_tokens: List[int] = []
        self.last_scores = None
        self.last_keep = None

    def get_prev_temporal_idx(self, layer_idx: int):
        if len(self.position_cache) <= layer_idx:
            return -1
        c = self.position_cache[layer_idx]
        return c[0, 0, -1] if c.ndim == 3 else c[0, -1]

    def _append_pos(self, pos, layer_idx):
        while len(self.position_cache) < layer_idx:
            self.position_cache.append([])
        if len(self.position_cache) == layer_idx:
            self.position_cache.append(pos)
        elif len(self.position_cache[layer_idx]) == 0:
            self.position_cache[layer_idx] = pos
        else:
            self.position_cache[layer_idx] = torch.cat([self.position_cache[layer_idx], pos], dim=-1)

    def update(self, key_states, value_states, layer_idx, query_states=None, position_ids=None,
               rotary_emb=None, mrope_section=None):
        while len(self.key_cache) <= layer_idx:
            self.key_cache.append(None)
            self.value_cache.append(None)
        if self.key_cache[layer_idx] is None:
            k_out, v_out = key_states, value_states
        else:
            k_out = torch.cat([self.key_cache[layer_idx], key_states], dim=2)
            v_out = torch.cat([self.value_cache[layer_idx], value_states], dim=2)
        self.key_cache[layer_idx], self.value_cache[layer_idx] = k_out, v_out
        if not self.kvcache_compression:
            if self.pos_embed_reforge:
                self._append_pos(position_ids, layer_idx)
            return k_out, v_out

        q_len = query_states.shape[2]
        k_len = key_states.shape[2]
        q, k = query_states, key_states
        if self.pos_embed_reforge:
            cos, sin = rotary_emb(value_states, position_ids)
            c1, s1 = select_mrope(cos, mrope_section), select_mrope(sin, mrope_section)
            sc = rotary_emb.attention_scaling
            q = unrotate(q, c1, s1, sc, self.semantics)
            k = unrotate(k, c1, s1, sc, self.semantics)
        keep_len = max(1, int(self.compression_ratio * q_len))
        score = pivot_scores(q, k, self.semantics)
        idx = keep_indices(score, keep_len, self.keypatches_mask_chunk, self.tie)
        self.last_scores, self.last_keep = score, idx
        kc = k[:, :, idx]
        vc = value_states[:, :, idx]
        pos_c = position_ids[..., idx].clone()
        if self.pos_embed_reforge:
            pos_c[0] = reforge_temporal(pos_c[0], keep_len, k_len)
            cos, sin = rotary_emb(vc, pos_c)
            kc = rotate(kc, select_mrope(cos, mrope_section), select_mrope(sin, mrope_section))
            self._append_pos(pos_c, layer_idx)
        while len(self.num_evicted_tokens) <= layer_idx:
            self.num_evicted_tokens.append(0)
        self.num_evicted_tokens[layer_idx] += k_len - keep_len
        self.key_cache[layer_idx] = torch.cat([k_out[..., :-q_len, :], kc], dim=2)
        self.value_cache[layer_idx] = torch.cat([v_out[..., :-q_len, :], vc], dim=2)
        return k_out, v_out
