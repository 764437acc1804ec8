"""Second oracle form: the SAME stock-torch op sequence the reference executes, device-agnostic.
TEST / BASELINE INFRASTRUCTURE ONLY.

``oracle/dpselect.py`` and ``oracle/pivotkv.py`` spell every rounding out; this module instead issues the library
calls the reference issues (``F.cosine_similarity``, ``F.max_pool1d_with_indices``, ``torch.topk``, bf16
``torch.matmul``, ``softmax(dtype=float32)`` ...; ``visual_compression.py:98-177``, ``longvideo_cache.py:244-318``),
so that (a) on CPU it is bit-identical to the frozen reference outputs and costs what the reference costs - it
is the ``cpu_baseline`` / ``--impl reference`` arm of bench.py - and (b) on the GPU box it is "the reference
executed with torch-CUDA ops", the parity target of the kernels.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def dpselect(bank: torch.Tensor, t: int, sync: bool):
    """-> (compressed [1, t, N, C], mask [t*N] bool, kept indices)."""
    _, T, N, C = bank.shape
    cos = F.cosine_similarity(bank[:, :-1], bank[:, 1:], dim=-1)[0]
    d = torch.cat([torch.ones_like(cos[:1], dtype=torch.float), 1 - cos.float()], dim=0)        # [T, N]
    rows = d.mean(1)[None] if sync else d.t().contiguous()                                      # [1 or N, T]
    arg = F.max_pool1d_with_indices(rows[:, None, :], 3, 1, padding=1)[1][:, 0]                 # [rows, T]
    is_peak = arg == torch.arange(T, device=bank.device)[None]
    keys = torch.where(is_peak, rows + 2, rows)
    kept = torch.topk(keys, k=t, sorted=False, dim=1)[1].sort(dim=1)[0]                          # [rows, t]
    if sync:
        idx = kept[0]
        return bank[:, idx], is_peak[0][idx][:, None].repeat(1, N).flatten(), idx
    idx = kept.t()
    out = bank.gather(1, idx[None, :, :, None].expand(-1, -1, -1, C))
    return out, is_peak.t().gather(0, idx).flatten(), idx


def _mrope_pick(tab, sections):
    if not sections:
        return tab.unsqueeze(1)
    parts = tab.split(list(sections) * 2, dim=-1)
    return torch.cat([p[i % 3] for i, p in enumerate(parts)], dim=-1).unsqueeze(1)


def _half_turn(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def pivot_update(q, k, v, ratio, keymask=None, position_ids=None, rotary=None, sections=None, reforge=False):
    """One compressing ``update`` without the cache bookkeeping: -> (kept K, kept V, kept positions, indices, score)."""
    _, H, L, D = q.shape
    KVH = k.shape[1]
    if reforge:
        cos, sin = rotary(v, position_ids)
        c, s = _mrope_pick(cos, sections), _mrope_pick(sin, sections)
        sc2 = rotary.attention_scaling ** 2
        q = ((q * c) - (_half_turn(q) * s)) / sc2
        k = ((k * c) - (_half_turn(k) * s)) / sc2
    kr = k[:, :, None].expand(1, KVH, H // KVH, L, D).reshape(1, H, L, D)
    w = torch.matmul(q, kr.transpose(2, 3)) / math.sqrt(D)
    w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
    score = w[0].sum(1).reshape(KVH, -1, L).mean(1).mean(0)
    if keymask is not None:
        score.masked_fill_(keymask, 1.0)
    keep = max(1, int(ratio * L))
    idx = score.topk(keep)[1].sort().values
    kk, vv = k[:, :, idx], v[:, :, idx]
    pos = position_ids[..., idx].clone() if position_ids is not None else None
    if reforge:
        lo = pos[0].min()
        pos[0] = lo + ((pos[0] - lo) * (keep / L)).long()
        cos, sin = rotary(vv, pos)
        c, s = _mrope_pick(cos, sections), _mrope_pick(sin, sections)
        kk = (kk * c) + (_half_turn(kk) * s)
    return kk, vv, pos, idx, score


def _closest_pair(bank, sync):
    """index [1, 1, N] of the most similar adjacent frame pair (``visual_compression.py:20-24``)"""
    cos = F.cosine_similarity(bank[:, :-1], bank[:, 1:], dim=-1)
    if sync:
        cos = cos.mean(-1, keepdim=True).expand(-1, -1, bank.shape[2])
    return torch.max(cos, dim=1, keepdim=True)[1]


def mallm_round(bank, size, sync=False):
    """one MA-LLM merge with stock torch ops: ``[1, T, N, C]``, ``[1, T, N]`` -> ``[1, T-1, N, C]``, ``[1, T-1, N]``"""
    _, T, N, C = bank.shape
    first = _closest_pair(bank, sync)
    rest = torch.arange(T - 1, device=bank.device)[None, :, None].repeat(1, 1, N)
    rest = rest + (rest > first).long()                                   # every frame but first + 1
    wide = lambda i: i.unsqueeze(-1).expand(-1, -1, -1, C)
    moved, moved_n = bank.gather(1, wide(first + 1)), size.gather(1, first + 1)
    kept, kept_n = bank.gather(1, wide(rest)), size.gather(1, rest)
    moved *= moved_n.unsqueeze(-1)
    kept *= kept_n.unsqueeze(-1)
    kept.scatter_add_(1, wide(first), moved)
    kept_n.scatter_add_(1, first, moved_n)
    return kept / kept_n.unsqueeze(-1), kept_n


def mallm_hard_round(bank, sync=False):
    """one MA-LLM-hard step: the first frame of the closest pair is overwritten by the second"""
    _, T, N, C = bank.shape
    first = _closest_pair(bank, sync)
    rest = torch.arange(T - 1, device=bank.device)[None, :, None].repeat(1, 1, N)
    rest = rest + (rest > first).long()
    wide = lambda i: i.unsqueeze(-1).expand(-1, -1, -1, C)
    kept = bank.gather(1, wide(rest))
    kept.scatter_(1, wide(first), bank.gather(1, wide(first + 1)))
    return kept


def mallm_compress(bank, t, sync=False, hard=False):
    """the caller's loop (``qwen2_vl.py:402-409``)"""
    size = torch.ones_like(bank[:, :, :, 0])
    while bank.shape[1] > t:
        if hard:
            bank = mallm_hard_round(bank, sync)
        else:
            bank, size = mallm_round(bank, size, sync)
    return (bank, None) if hard else (bank, size)
