"""Splitting ONE video across the GPUs of a box.  Whole videos need no code: one process per
GPU, each running the single-GPU operators (``bench.py --gpus N``; the reference's own pattern, ``infer_eval.py:181``).

Two exchange steps exist inside a video, both tiny all-gathers over NCCL (or gloo in the CPU tests of the host logic):

* **DPSelect by frame range**: rank r owns frames ``[t0, t1)`` plus the last frame of rank r-1 as a halo, computes its
  rows of ``dis`` (``rtk_dpselect_dis(halo=1)``), all ranks all-gather ``dis`` (fp32 ``[T, N]``, <= 6 MB) because peaks
  look one frame across the seam and the top-t is global, run the identical selection, and compact only the survivors
  they own.  Kept indices are bit-identical on every rank by construction.
* **PivotKV by KV head**: rank r owns KV heads ``[g0, g1)`` and the query heads of those groups, computes their
  line-269 score rows (``rtk_pivot_score``), all ranks all-gather the ``[KVH, L]`` bf16 rows (32-50 KB), take the mean
  over KV heads in the reference's order, select identically, and compact their own heads.

The partition / gather helpers are device-agnostic torch code so that they can be tested with gloo on CPU; only the
functions that take CUDA tensors call into the library.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------ partition helpers
def split_range(n: int, world: int) -> List[Tuple[int, int]]:
    """contiguous, balanced: the first ``n % world`` parts get one extra element"""
    base, extra = divmod(n, world)
    out, a = [], 0
    for r in range(world):
        b = a + base + (1 if r < extra else 0)
        out.append((a, b))
        a = b
    return out


def all_gather_rows(local: torch.Tensor, sizes: List[int], group=None) -> torch.Tensor:
    """concatenate per-rank tensors with different first dims (``sizes[r]`` rows on rank r) along dim 0"""
    world = dist.get_world_size(group)
    assert len(sizes) == world and local.shape[0] == sizes[dist.get_rank(group)]
    mx = max(sizes)
    if mx == 0:
        return local.new_empty((0,) + tuple(local.shape[1:]))
    pad = local
    if local.shape[0] < mx:
        pad = local.new_zeros((mx,) + tuple(local.shape[1:]))
        pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad.contiguous(), group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def owned_slots(idx: torch.Tensor, t0: int, t1: int, n_patches: int):
    """Which output slots of the compacted ``[t, N]`` grid come from frames ``[t0, t1)``.

    ``idx`` is ``[t, N]`` (per-patch mode) or ``[t]`` (sync mode).  Returns (slots, frames): flat slot numbers
    ``j * N + p`` in ascending order and the source frame of each."""
    full = idx if idx.dim() == 2 else idx[:, None].expand(-1, n_patches)
    own = (full >= t0) & (full < t1)
    slots = torch.nonzero(own.reshape(-1))[:, 0]
    return slots, full.reshape(-1)[slots].long()


# --------------------------------------------------------------------------------------------- DPSelect shard
def dpselect_frame_sharded(x_local: torch.Tensor, t0: int, t1: int, T: int, tgt_mem_len: int, sync: bool = False,
                           group=None):
    """DPSelect of one video whose frames are split across ranks.

    ``x_local``: bf16 ``[t1 - t0 (+1), N, C]`` on this rank's GPU - frames ``[t0, t1)`` preceded by frame ``t0 - 1``
    when ``t0 > 0``.  Returns ``(rows, slots, mask, idx)``: this rank's surviving embeddings ``[n_own, C]`` in slot
    order, their flat slot numbers in the global ``[t, N]`` output, and the replicated key-patch mask ``[t*N]`` and kept
    indices.  ``assemble_compacted`` rebuilds the reference's ``[1, t, N, C]`` tensor when one rank needs it whole."""
    from . import _native as N_
    from . import visual_compression as vc
    world = dist.get_world_size(group)
    halo = t0 > 0
    assert x_local.shape[0] == (t1 - t0) + int(halo)
    n_patches, C = x_local.shape[1], x_local.shape[2]
    dis_local = vc.dpselect_distance(x_local, halo=halo)                       # [t1 - t0, N]
    sizes = [b - a for a, b in split_range(T, world)]
    assert sizes[dist.get_rank(group)] == t1 - t0, "frame ranges must come from split_range(T, world)"
    dis = all_gather_rows(dis_local, sizes, group)                             # [T, N], identical everywhere
    idx, mask = vc.dpselect_select(dis, tgt_mem_len, sync)
    slots, frames = owned_slots(idx, t0, t1, n_patches)
    src_row = (frames - (t0 - int(halo))) * n_patches + slots % n_patches     # row of x_local.view(-1, C)
    rows = torch.empty((slots.numel(), C), dtype=x_local.dtype, device=x_local.device)
    xl = x_local.contiguous()
    with torch.cuda.device(xl.device):
        N_.check(N_.lib().rtk_gather_rows(xl.data_ptr(), C * xl.element_size(), src_row.data_ptr(), slots.numel(),
                                          rows.data_ptr(), N_.stream_ptr(xl.device)), "rtk_gather_rows")
    return rows, slots, mask, idx


def assemble_compacted(rows: torch.Tensor, slots: torch.Tensor, t: int, n_patches: int, group=None) -> torch.Tensor:
    """all-gather every rank's surviving rows into the reference's ``[1, t, N, C]`` layout (slot = j * N + p)"""
    world = dist.get_world_size(group)
    cnt = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    cnts = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    sizes = [int(c) for c in cnts]
    all_rows = all_gather_rows(rows, sizes, group)
    all_slots = all_gather_rows(slots, sizes, group)
    out = rows.new_empty((t * n_patches, rows.shape[1]))
    out[all_slots] = all_rows
    return out.reshape(1, t, n_patches, rows.shape[1])


# ---------------------------------------------------------------------------------------------- PivotKV shard
def gather_head_scores(head_scores_local: torch.Tensor, kv_heads_per_rank: List[int], group=None) -> torch.Tensor:
    """``[KVH_local, L]`` -> ``[KVH, L]`` in global KV-head order (rank r owns a contiguous block of heads)"""
    return all_gather_rows(head_scores_local, kv_heads_per_rank, group)


def pivot_update_kv_sharded(query_local, key_local, value_local, keep_len: int, kv_heads_per_rank: List[int],
                            keymask: Optional[torch.Tensor] = None, position_ids: Optional[torch.Tensor] = None,
                            rotary_emb=None, mrope_section=None, reforge: bool = False, group=None):
    """One compressing update with the KV heads of the chunk split across ranks.

    ``query_local [1, G * KVH_local, L, D]`` are the query heads of this rank's KV groups.  Returns
    ``(kept_k [1, KVH_local, keep, D], kept_v, kept_positions, keep_idx, head_scores [KVH, L])``; ``keep_idx`` is
    identical on every rank."""
    from . import longvideo_cache as lc
    q, k = query_local, key_local
    if reforge:
        fast = lc._rotary_inv_freq(rotary_emb)
        scaling = float(rotary_emb.attention_scaling)
        D = q.shape[-1]
        if fast is not None:
            cos, sin = lc.pivot_rope_tables(position_ids, fast[0], D, mrope_section, scaling)
            sec = None
        else:
            cos, sin = rotary_emb(value_local, position_ids)
            sec = mrope_section
        q = lc.pivot_rope(q, cos, sin, sec, scaling, forward=False)
        k = lc.pivot_rope(k, cos, sin, sec, scaling, forward=False)
    hs_local = lc.pivot_head_scores(q, k)
    hs = gather_head_scores(hs_local, kv_heads_per_rank, group)
    keep_idx = lc.pivot_select(hs, keep_len, keymask)
    kept_k, kept_v, kept_pos = lc.pivot_compact(k, value_local, keep_idx, position_ids, reforge=reforge)
    if reforge:
        if fast is not None:
            cos2, sin2 = lc.pivot_rope_tables(kept_pos, fast[0], q.shape[-1], mrope_section, scaling)
            lc.pivot_rope(kept_k, cos2, sin2, None, 1.0, forward=True, out=kept_k)
        else:
            cos2, sin2 = rotary_emb(kept_v, kept_pos)
            lc.pivot_rope(kept_k, cos2, sin2, mrope_section, 1.0, forward=True, out=kept_k)
    return kept_k, kept_v, kept_pos, keep_idx, hs
