"""Oracle for the MA-LLM memory-bank compressors.  TEST INFRASTRUCTURE ONLY.

Restates ``retake/visual_compression.py:5-47`` (``memory_bank_compress_MALLM``: merge the most similar adjacent
frame pair, weighted by how many frames each side already stands for) and ``:50-83``
(``memory_bank_compress_MALLM_hard``: drop the first frame of that pair) with every rounding written out, plus the
caller's loop (``qwen2_vl.py:402-409``, ``llava_onevision.py:235-243``: repeat until ``tgt_mem_len`` frames are left).

What one round of the reference does to a bank ``x[T, N, C]`` with sizes ``s[T, N]`` (batch 1), per patch column p:

* ``sim[i] = cosine(x[i], x[i+1])`` in the bank dtype (``oracle.dpselect.adjacent_cosine_similarity``);
  ``sync=True`` replaces it by its mean over the patches (``Tensor.mean(-1)`` in the bank dtype).
* ``m = argmax_i sim[i]`` - ``torch.max(dim=1)``; among equal maxima the LOWEST index (``tie="lowest"``: ATen's
  CPU and CUDA reductions both break ties towards the smaller index; NaN counts as the maximum).
* every surviving row is multiplied by its size and divided by it again, each op rounded to the bank dtype
  (``:38-39,46``): ``x'[i] = rd(rd(x[i] * s[i]) / s[i])`` - NOT the identity in bf16 when ``s`` is no power of two;
* row ``m`` becomes ``rd(rd(rd(x[m] * s[m]) + rd(x[m+1] * s[m+1])) / rd(s[m] + s[m+1]))`` (``:42-46``), row
  ``m+1`` disappears, ``s[m] = rd(s[m] + s[m+1])`` (sizes live in the bank dtype too: bf16 stops counting at 256).

``rd`` is round-to-nearest-even to the bank dtype (identity for fp32 banks).  The hard variant has no arithmetic:
position m takes the content of m+1 and m+1 is dropped, i.e. frame m is deleted.
"""
from __future__ import annotations

import torch

from .dpselect import BF16, _r, adjacent_cosine_similarity


def _rd(x: torch.Tensor, lowp: bool) -> torch.Tensor:
    return _r(x) if lowp else x


def aten_cuda_rowsum_general(v: torch.Tensor, vec: int, elem_bytes: int, base_offset_bytes: int = 0) -> torch.Tensor:
    """fp32 sums over the last dim of ``v[R, n]`` (fp32 container, rows contiguous in a tensor of ``elem_bytes``-wide
    elements starting ``base_offset_bytes`` past a 16-byte boundary) in the order of ATen's CUDA reduction
    (Reduce.cuh): n >= 128 -> ``min(last_pow2(n // vec), 32)`` lanes, ``vec``-element vectors with ATen's
    unaligned-head / scalar-tail handling; n < 128 -> block width ``min(last_pow2(n), 32)``, lane l owns l, l + w, ... dealt to four accumulators."""
    R, n = v.shape
    vv = v.to(torch.float32).cpu()
    out = torch.empty(R, dtype=torch.float32)
    rows = torch.arange(R)
    if n >= 128:
        shifts = ((base_offset_bytes + rows * n * elem_bytes) % (vec * elem_bytes)) // elem_bytes
    else:
        shifts = torch.zeros(R, dtype=torch.long)
    for shift in shifts.unique().tolist():
        sel = rows[shifts == shift]
        x = vv[sel]
        acc = torch.zeros(len(sel), 32, max(vec, 4))
        width = 32
        if n >= 128:
            start = 0
            if shift > 0:                                    # head: lanes shift..vec-1 take one element each
                head = vec - shift
                acc[:, shift:vec, 0] = acc[:, shift:vec, 0] + x[:, :head]
                start = head
            body = (n - start) // vec
            width = 1                                        # block width: last_pow2(n // vec), at most a warp
            while width * 2 <= n // vec and width < 32:
                width *= 2
            k = 0
            while k * width < body:
                lanes = min(width, body - k * width)
                chunk = x[:, start + k * width * vec: start + (k * width + lanes) * vec].reshape(len(sel), lanes, vec)
                acc[:, :lanes, :vec] = acc[:, :lanes, :vec] + chunk
                k += 1
            tail = x[:, start + body * vec:]                 # < vec elements, lane j takes element j
            acc[:, :tail.shape[1], 0] = acc[:, :tail.shape[1], 0] + tail
            nacc = vec
        else:
            width = 1
            while width * 2 <= n and width < 32:
                width *= 2
            for i0 in range(0, n, width):
                k = (i0 // width) % 4
                lanes = min(width, n - i0)
                acc[:, :lanes, k] = acc[:, :lanes, k] + x[:, i0:i0 + lanes]
            nacc = 4
        lane_sum = acc[:, :, 0].clone()
        for j in range(1, nacc):
            lane_sum = lane_sum + acc[:, :, j]
        off = width >> 1
        while off:
            lane_sum[:, :off] = lane_sum[:, :off] + lane_sum[:, off:2 * off]
            off >>= 1
        out[sel] = lane_sum[:, 0]
    return out


def aten_cuda_bf16_rowmean(v: torch.Tensor, vec: int = 8) -> torch.Tensor:
    """``bf16(sum * fp32(1/n))`` over the last dim of a contiguous bf16 ``[R, n]`` tensor, ATen-CUDA order"""
    n = v.shape[1]
    s = aten_cuda_rowsum_general(v, vec, 2)
    return _r(s * (torch.tensor(1.0, dtype=torch.float32) / torch.tensor(float(n), dtype=torch.float32)))


def first_argmax(sim: torch.Tensor) -> torch.Tensor:
    """index of the maximum along dim 0 of ``sim[T-1, ...]``; lowest index among equals, NaN is the maximum"""
    key = torch.where(torch.isnan(sim), torch.full_like(sim, float("inf")), sim)
    isn = torch.isnan(sim).any(0)
    best = key.max(0).values
    hit = (key == best[None]) & (torch.isnan(sim) | ~isn[None])
    T1 = sim.shape[0]
    ar = torch.arange(T1, device=sim.device).reshape((T1,) + (1,) * (sim.dim() - 1)).expand_as(sim)
    return torch.where(hit, ar, torch.full_like(ar, T1)).min(0).values


def similarity(bank: torch.Tensor, sync: bool, reduce: str = "torch", mean_vec: int = 8) -> torch.Tensor:
    """``[T-1, N]`` similarity the argmax looks at (``visual_compression.py:20-22``)"""
    sim = adjacent_cosine_similarity(bank, reduce)
    if sync:
        if reduce == "aten_cuda" and bank.dtype == BF16:
            mean = aten_cuda_bf16_rowmean(sim, mean_vec).to(sim.device)
        else:
            mean = sim.to(bank.dtype).mean(-1).to(torch.float32)
        sim = mean[:, None].expand(-1, sim.shape[1])
    return sim


def mallm_round(bank: torch.Tensor, size: torch.Tensor, sync: bool = False, reduce: str = "torch", mean_vec: int = 8):
    """one merge: ``bank[T, N, C]``, ``size[T, N]`` -> ``([T-1, N, C], [T-1, N], merged index [N])``"""
    T, N, C = bank.shape
    lowp = bank.dtype == BF16
    m = first_argmax(similarity(bank, sync, reduce, mean_vec))                 # [N]
    xf, sf = bank.to(torch.float32), size.to(torch.float32)
    scaled = _rd(xf * sf[..., None], lowp)                                     # every row times its size
    cols = torch.arange(N, device=bank.device)
    new_size = sf.clone()
    new_size[m, cols] = _rd(sf[m, cols] + sf[m + 1, cols], lowp)
    scaled[m, cols] = _rd(scaled[m, cols] + scaled[m + 1, cols], lowp)
    keep = torch.ones(T, N, dtype=torch.bool, device=bank.device)
    keep[m + 1, cols] = False
    order = torch.sort((~keep).to(torch.int8), dim=0, stable=True).indices[:T - 1]    # surviving frames, ascending
    scaled = scaled.gather(0, order[..., None].expand(-1, -1, C))
    new_size = new_size.gather(0, order)
    out = _rd(scaled / new_size[..., None], lowp)
    return out.to(bank.dtype), new_size.to(size.dtype), m


def mallm_hard_round(bank: torch.Tensor, sync: bool = False, reduce: str = "torch", mean_vec: int = 8):
    """one deletion: ``bank[T, N, C]`` -> ``([T-1, N, C], deleted index [N])``"""
    T, N, C = bank.shape
    m = first_argmax(similarity(bank, sync, reduce, mean_vec))
    keep = torch.ones(T, N, dtype=torch.bool, device=bank.device)
    keep[m, torch.arange(N, device=bank.device)] = False
    order = torch.sort((~keep).to(torch.int8), dim=0, stable=True).indices[:T - 1]
    return bank.gather(0, order[..., None].expand(-1, -1, C)), m


def mallm_compress(bank: torch.Tensor, tgt_mem_len: int, sync: bool = False, hard: bool = False, reduce: str = "torch",
                   mean_vec: int = 8):
    """the caller's loop (``qwen2_vl.py:402-409``): ``bank[1, T, N, C]`` -> ``[1, t, N, C]`` (and sizes ``[1, t, N]``)"""
    x = bank[0]
    size = torch.ones_like(x[:, :, 0])
    while x.shape[0] > tgt_mem_len:
        if hard:
            x, _ = mallm_hard_round(x, sync, reduce, mean_vec)
        else:
            x, size, _ = mallm_round(x, size, sync, reduce, mean_vec)
    return (x[None], None) if hard else (x[None], size[None])
