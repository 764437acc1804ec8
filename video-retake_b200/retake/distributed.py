"""Splitting ONE video across the GPUs of a box.  Whole videos need no code: one process per
GPU, each running the single-GPU operators (``bench.py --gpus N``; the reference's own pattern, ``infer_eval.py:181``).

Two exchange steps exist inside a video:

* **DPSelect by frame range**: rank r owns frames ``[t0, t1)`` plus the last frame of rank r-1 as a halo, computes its
  rows of ``dis`` (``rtk_dpselect_dis(halo=1)``), all ranks gather ``dis`` (fp32 ``[T, N]``, <= 6 MB; ONE
  ``all_gather_into_tensor``) because peaks look one frame across the seam and the top-t is global, run the identical
  selection, and ``rtk_dpselect_gather_owned`` writes the survivors a rank owns straight into their output slots
  (``dpselect_frame_sharded_fused``; ``dpselect_frame_sharded`` is the round-1 form that returns rows + slot numbers).
  Kept indices are bit-identical on every rank by construction.
* **PivotKV by KV head**: rank r owns KV heads ``[g0, g1)`` and the query heads of those groups.  Two C-ABI calls around
  one exchange of the ``[KVH_local, L]`` bf16 score rows (16-50 KB): ``rtk_pivot_update(skip_select)`` un-rotates and
  scores the local heads, the rows travel as NVLink peer stores issued by the library's own put kernel (``ScoreExchange``,
  torch symmetric memory) or as one NCCL ``all_gather_into_tensor``, and ``rtk_pivot_update(skip_score)`` takes the mean
  over ALL KV heads in the reference's order, selects identically on every rank and compacts the local heads
  (``pivot_update_kv_sharded``; ``pivot_update_batch_kv_sharded`` does it for all layers of a chunk with one exchange).

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


# ----------------------------------------------------------------------------- DPSelect shard, sync-free form
_GATHER_BUF = {}


def _gather_buffer(dev, n, dtype, tag):
    key = (dev, dtype, tag)
    buf = _GATHER_BUF.get(key)
    if buf is None or buf.numel() < n:
        buf = _GATHER_BUF[key] = torch.empty(n, dtype=dtype, device=dev)
    return buf[:n]


def dpselect_frame_sharded_fused(x_local: torch.Tensor, t0: int, t1: int, T: int, tgt_mem_len: int, sync: bool = False,
                                 group=None, out: Optional[torch.Tensor] = None, zero_fill: bool = False):
    """The frame-range split as four device operations and NO host synchronisation: local distances
    (``rtk_dpselect_dis(halo)``), ONE ``all_gather_into_tensor`` of the ``[T, N]`` fp32 distances into a preallocated
    buffer, the replicated selection, and ``rtk_dpselect_gather_owned``, which writes this rank's surviving rows straight
    into their slots of the reference's ``[1, t, N, C]`` layout.

    Returns ``(out, mask, idx)``: ``out`` (``[1, t, N, C]``, the caller's buffer or a fresh one) has only the slots whose
    source frame lies in ``[t0, t1)`` written - the rest is left as it was (zeros with ``zero_fill``); mask and indices are
    identical on every rank.  ``assemble_owned`` sums zero-filled partial outputs when one rank needs the whole tensor."""
    from . import _native as N_
    from . import visual_compression as vc
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    halo = t0 > 0
    assert x_local.shape[0] == (t1 - t0) + int(halo)
    n_patches, C = x_local.shape[1], x_local.shape[2]
    ranges = split_range(T, world)
    assert ranges[rank] == (t0, t1), "frame ranges must come from split_range(T, world)"
    dis_local = vc.dpselect_distance(x_local, halo=halo)                       # [t1 - t0, N]
    mx = max(b - a for a, b in ranges)
    dev = x_local.device
    if all(b - a == mx for a, b in ranges):
        dis = _gather_buffer(dev, T * n_patches, torch.float32, "dis").view(T, n_patches)
        dist.all_gather_into_tensor(dis, dis_local, group=group)
    else:
        # ragged split: equal-sized padded blocks through the same single collective, then one compaction copy
        padded = _gather_buffer(dev, world * mx * n_patches, torch.float32, "dis_pad").view(world, mx, n_patches)
        mine = _gather_buffer(dev, mx * n_patches, torch.float32, "dis_mine").view(mx, n_patches)
        mine[: t1 - t0].copy_(dis_local)
        dist.all_gather_into_tensor(padded, mine, group=group)
        dis = torch.cat([padded[r, : b - a] for r, (a, b) in enumerate(ranges)], dim=0)
    idx, mask = vc.dpselect_select(dis, tgt_mem_len, sync)
    if out is None:
        alloc = torch.zeros if zero_fill else torch.empty
        out = alloc((1, tgt_mem_len, n_patches, C), dtype=x_local.dtype, device=dev)
    xl = x_local.contiguous()
    with torch.cuda.device(dev):
        N_.check(N_.lib().rtk_dpselect_gather_owned(xl.data_ptr(), xl.shape[0], t0 - int(halo), t0, t1, n_patches, C,
                                                    idx.data_ptr(), tgt_mem_len, int(sync), out.data_ptr(),
                                                    N_.stream_ptr(dev)), "rtk_dpselect_gather_owned")
    return out, mask, idx


def assemble_owned(out_partial: torch.Tensor, group=None) -> torch.Tensor:
    """every slot is filled on exactly one rank and zero elsewhere: a sum of the bit patterns rebuilds the whole tensor"""
    bits = out_partial.contiguous().view(torch.int32).clone()      # (NCCL has no 16-bit integer sum; C is even)
    dist.all_reduce(bits, op=dist.ReduceOp.SUM, group=group)
    return bits.view(out_partial.dtype)


# ---------------------------------------------------------------------------------------------- PivotKV shard
def gather_head_scores(head_scores_local: torch.Tensor, kv_heads_per_rank: List[int], group=None) -> torch.Tensor:
    """``[KVH_local, L]`` -> ``[KVH, L]`` in global KV-head order (rank r owns a contiguous block of heads)"""
    return all_gather_rows(head_scores_local, kv_heads_per_rank, group)


class ScoreExchange:
    """Where the per-KV-head score rows of every rank meet (one per process group and device).

    ``p2p``: two ``[rows, L]`` bf16 buffers and two sets of flag words in torch symmetric memory, mapped into every rank's
    address space; the library's one-CTA put kernel stores a rank's rows into all peers over NVLink and raises a flag, the
    select kernel of every rank waits for all flags (``rtk_pivot_update_args.xchg_*``) - no collective launch at all.
    ``nccl``: the same two calls around ONE ``all_gather_into_tensor`` on a preallocated buffer (used when symmetric memory
    is not available, or when asked for with RTK_SHARD_TRANSPORT=nccl)."""

    _cache = {}

    @classmethod
    def get(cls, group, device, rows: int, L: int, transport: Optional[str] = None):
        import os
        transport = transport or os.environ.get("RTK_SHARD_TRANSPORT", "p2p")
        # keyed by the ranks of the group (an id() could be recycled by another group object)
        key = (tuple(dist.get_process_group_ranks(group if group is not None else dist.group.WORLD)), device, transport)
        ex = cls._cache.get(key)
        if ex is None or ex.capacity < rows * L:
            ex = cls._cache[key] = cls(group, device, max(rows * L, 8 * 16384), transport)
        return ex

    def __init__(self, group, device, capacity: int, transport: str):
        self.group, self.device, self.capacity = group, device, capacity
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.epoch = 0
        self.transport = transport
        self.handles = None
        if transport == "p2p":
            try:
                import torch.distributed._symmetric_memory as sm
                g = group if group is not None else dist.group.WORLD
                self.scores = sm.empty((2, capacity), dtype=torch.bfloat16, device=device)
                self.flags = sm.empty((2, 8), dtype=torch.int32, device=device)
                self.flags.zero_()
                hs, hf = sm.rendezvous(self.scores, g), sm.rendezvous(self.flags, g)
                self.score_ptrs = [int(p) for p in hs.buffer_ptrs]
                self.flag_ptrs = [int(p) for p in hf.buffer_ptrs]
                self.handles = (hs, hf)
                torch.cuda.synchronize(device)
                dist.barrier(group)                       # nobody raises a flag before everybody has cleared theirs
            except Exception as e:  # noqa: BLE001 - symmetric memory is optional: NCCL carries the same rows
                import warnings
                warnings.warn(f"symmetric memory unavailable ({e!r}); the KV-head split uses NCCL all_gather_into_tensor")
                self.transport = "nccl"
        if self.transport != "p2p":
            self.scores = torch.empty((2, capacity), dtype=torch.bfloat16, device=device)

    def begin(self):
        """next exchange: (parity, epoch)"""
        self.epoch += 1
        return self.epoch & 1, self.epoch

    def own_ptr(self, parity: int) -> int:
        return self.scores.data_ptr() + parity * self.capacity * 2


def pivot_update_kv_sharded(query_local, key_local, value_local, keep_len: int, kv_heads_per_rank: List[int],
                            keymask: Optional[torch.Tensor] = None, position_ids: Optional[torch.Tensor] = None,
                            rotary_emb=None, mrope_section=None, reforge: bool = False, group=None,
                            transport: Optional[str] = None):
    """One compressing update with the KV heads of the chunk split across ranks.

    ``query_local [1, G * KVH_local, L, D]`` are the query heads of this rank's KV groups.  Returns
    ``(kept_k [1, KVH_local, keep, D], kept_v, kept_positions, keep_idx, head_scores [KVH, L])``; ``keep_idx`` is
    identical on every rank.  ``head_scores`` is a VIEW of the exchange buffer: valid until the next update but one on
    this group (clone it to keep it).

    Two C-ABI calls around one exchange of the ``[KVH_local, L]`` score rows (``rtk_pivot_update`` with ``skip_select``,
    then with ``skip_score``): un-rotate + score + put | wait + select + compact + re-rotate, eight launches, the rows
    travelling as NVLink peer stores issued by the library's own kernel (``ScoreExchange``), or as one NCCL
    ``all_gather_into_tensor``.  Ranks with different numbers of heads take the unfused chain below."""
    if len(set(kv_heads_per_rank)) != 1 or query_local.dtype != torch.bfloat16:
        return _pivot_update_kv_sharded_unfused(query_local, key_local, value_local, keep_len, kv_heads_per_rank, keymask,
                                                position_ids, rotary_emb, mrope_section, reforge, group)
    import ctypes as C

    from . import _native as N_
    from . import longvideo_cache as lc
    dev = query_local.device
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    kvh_local = key_local.shape[1]
    rows, L = kvh_local * world, query_local.shape[2]
    ex = ScoreExchange.get(group, dev, rows, L, transport)
    parity, epoch = ex.begin()
    a, outs, keepalive, fast, _ = lc.fill_update_args(keymask, reforge, lc._inv_freq_on_device, query_local, key_local, value_local,
                                                      position_ids, rotary_emb, mrope_section, keep_len)
    lib = N_.lib()
    ws = lc._workspace(dev, int(lib.rtk_pivot_update_workspace_bytes(a.H, a.KVH, a.L, a.D)) + 256)
    ws_ptr = (ws.data_ptr() + 255) & ~255
    a.workspace, a.workspace_bytes = ws_ptr, ws.numel() - (ws_ptr - ws.data_ptr())
    own = ex.own_ptr(parity)
    hs_all = ex.scores[parity, : rows * L].view(rows, L)
    a.head_scores = own + rank * kvh_local * L * 2
    a.skip_select, a.skip_score, a.score_rows = 1, 0, rows
    if ex.transport == "p2p":
        a.xchg_world, a.xchg_rank, a.xchg_epoch = world, rank, epoch
        for r in range(world):
            a.xchg_scores[r] = ex.score_ptrs[r] + parity * ex.capacity * 2
            a.xchg_flags[r] = ex.flag_ptrs[r] + parity * 8 * 4
    with torch.cuda.device(dev):
        N_.check(lib.rtk_pivot_update(C.byref(a), N_.stream_ptr(dev)), "rtk_pivot_update(skip_select)")
        if ex.transport != "p2p":
            dist.all_gather_into_tensor(hs_all, hs_all[rank * kvh_local:(rank + 1) * kvh_local], group=group)
        a.skip_select, a.skip_score = 0, 1
        a.head_scores = own
        N_.check(lib.rtk_pivot_update(C.byref(a), N_.stream_ptr(dev)), "rtk_pivot_update(skip_score)")
    kept_k, kept_v, kept_pos = outs["k_out"], outs["v_out"], outs["pos_out"]
    if reforge and not fast:
        cos2, sin2 = rotary_emb(kept_v, kept_pos)
        lc.pivot_rope(kept_k, cos2, sin2, mrope_section, 1.0, forward=True, out=kept_k)
    del keepalive
    return kept_k, kept_v, kept_pos, outs["keep_idx"], hs_all


def pivot_update_batch_kv_sharded(layers, keep_len: int, kv_heads_per_rank: List[int], rotary_emb=None, mrope_section=None,
                                  reforge: bool = False, group=None, transport: Optional[str] = None):
    """The KV-head split for ALL layers of a chunk at once (deferred compression, SURVEY.md 8(f2) + 8(e)): two
    ``rtk_pivot_update_batch`` calls around ONE exchange of every layer's ``[KVH_local, L]`` score rows.

    ``layers``: list of ``(query_local, key_local, value_local, keymask or None, position_ids or None)`` with equal shapes.
    Returns a list of ``(kept_k, kept_v, kept_positions, keep_idx)`` per layer; ``keep_idx`` is identical on every rank.
    Needs equal heads per rank and a rotary module with a static ``inv_freq`` when ``reforge`` (the batched kernels'
    envelope); at most 32 layers per call."""
    import ctypes as C

    from . import _native as N_
    from . import longvideo_cache as lc
    assert len(set(kv_heads_per_rank)) == 1 and 1 <= len(layers) <= 32
    dev = layers[0][0].device
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    kvh_local, L = layers[0][1].shape[1], layers[0][0].shape[2]
    rows, n = kvh_local * world, len(layers)
    ex = ScoreExchange.get(group, dev, rows * n, L, transport)
    parity, epoch = ex.begin()
    arr = (lc._UpdateArgs * n)()
    outs_all, keep = [], []
    for i, (ql, kl, vl, km, pos) in enumerate(layers):
        a, outs, keepalive, fast, _ = lc.fill_update_args(km, reforge, lc._inv_freq_on_device, ql, kl, vl, pos, rotary_emb,
                                                          mrope_section, keep_len)
        if reforge and not fast:
            raise ValueError("the batched split needs a rotary module with a static inv_freq")
        base = parity * ex.capacity * 2 + i * rows * L * 2
        a.head_scores = ex.scores.data_ptr() + base + rank * kvh_local * L * 2
        a.skip_select, a.skip_score, a.score_rows = 1, 0, rows
        if ex.transport == "p2p":
            a.xchg_world, a.xchg_rank, a.xchg_epoch = world, rank, epoch
            for r in range(world):
                a.xchg_scores[r] = ex.score_ptrs[r] + base
                a.xchg_flags[r] = ex.flag_ptrs[r] + parity * 8 * 4
        arr[i] = a
        outs_all.append(outs)
        keep.append(keepalive)
    lib = N_.lib()
    a0 = arr[0]
    ws = lc._workspace(dev, int(lib.rtk_pivot_update_batch_workspace_bytes(a0.H, a0.KVH, a0.L, a0.D, n)) + 256)
    ws_ptr = (ws.data_ptr() + 255) & ~255
    ws_bytes = ws.numel() - (ws_ptr - ws.data_ptr())
    with torch.cuda.device(dev):
        N_.check(lib.rtk_pivot_update_batch(arr, n, ws_ptr, ws_bytes, N_.stream_ptr(dev)), "rtk_pivot_update_batch(skip_select)")
        if ex.transport != "p2p":
            blk = ex.scores[parity, : n * rows * L].view(n, rows, L)
            with dist._coalescing_manager(group=group, device=dev, async_ops=False):
                for i in range(n):
                    dist.all_gather_into_tensor(blk[i], blk[i, rank * kvh_local:(rank + 1) * kvh_local], group=group)
        for i in range(n):
            arr[i].skip_select, arr[i].skip_score = 0, 1
            arr[i].head_scores = ex.scores.data_ptr() + parity * ex.capacity * 2 + i * rows * L * 2
        N_.check(lib.rtk_pivot_update_batch(arr, n, ws_ptr, ws_bytes, N_.stream_ptr(dev)), "rtk_pivot_update_batch(skip_score)")
    del keep
    return [(o["k_out"], o["v_out"], o["pos_out"], o["keep_idx"]) for o in outs_all]


def _pivot_update_kv_sharded_unfused(query_local, key_local, value_local, keep_len: int, kv_heads_per_rank: List[int],
                                     keymask: Optional[torch.Tensor] = None, position_ids: Optional[torch.Tensor] = None,
                                     rotary_emb=None, mrope_section=None, reforge: bool = False, group=None):
    """the same update as a chain of the unfused entry points (any split of the heads over the ranks)"""
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
