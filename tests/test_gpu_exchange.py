"""GPU (one device): the KV-head split's two-call protocol (``rtk_pivot_update`` with ``skip_select`` / ``skip_score`` and the
``xchg_*`` peer exchange, ABI 3) with BOTH ranks emulated by one process on one stream - every "peer" buffer is a local
tensor, so the put kernel's stores, the flag words and the select kernel's wait run exactly as they do over NVLink
(tests/test_gpu_distributed.py runs the real 2-GPU version; it is skipped on a one-GPU box).  Results must equal the
single-GPU fused update bit for bit; the sync-free DPSelect compaction (``rtk_dpselect_gather_owned``) is checked the
same way."""
import ctypes as C

import pytest
import torch

from helpers import TableRotary
from test_gpu_pivotkv import qkv

pytestmark = pytest.mark.gpu


def _lc():
    from retake import longvideo_cache as lc
    return lc


def _call(a):
    from retake import _native as N
    N.check(N.lib().rtk_pivot_update(C.byref(a), N.stream_ptr(torch.device("cuda"))), "rtk_pivot_update")


@pytest.mark.parametrize("reforge", [False, True])
@pytest.mark.parametrize("L", [1024, 333])
def test_two_emulated_ranks_equal_single_gpu(L, reforge):
    lc = _lc()
    from retake import _native as N
    H, KVH, D, world, keep, mrope = 28, 4, 128, 2, max(1, L // 4), [16, 24, 24]
    G, per = H // KVH, KVH // world
    rot = TableRotary(D)
    rot.inv_freq = rot.inv_freq.cuda()
    ar = torch.arange(L, device="cuda")
    pos = torch.stack([5 + ar // 64, (ar % 64) // 8, ar % 8])[:, None]
    bufs = [torch.zeros(2, KVH * L, dtype=torch.bfloat16, device="cuda") for _ in range(world)]      # [parity][rows * L] per rank
    flags = [torch.zeros(2, 8, dtype=torch.int32, device="cuda") for _ in range(world)]
    lib = N.lib()
    for epoch in range(1, 5):                                # four exchanges: both parities twice, flags re-used
        parity = epoch & 1
        q, k, v = qkv(H, KVH, L, D, 3.0, seed=40 + epoch)
        mask = (torch.rand(L, generator=torch.Generator().manual_seed(epoch)) < 0.2).cuda()
        want_k, want_v, want_p, want_idx, want_hs = lc.pivot_update(q, k, v, keep, mask, pos, rot, mrope, reforge)
        args, outs = [], []
        ws = []
        for r in range(world):
            ql, kl, vl = q[:, r * per * G:(r + 1) * per * G], k[:, r * per:(r + 1) * per], v[:, r * per:(r + 1) * per]
            a, o, keepalive, fast, _ = lc.fill_update_args(mask, reforge, lc._inv_freq_on_device, ql, kl, vl, pos, rot, mrope, keep)
            assert fast
            w = torch.empty(int(lib.rtk_pivot_update_workspace_bytes(a.H, a.KVH, a.L, a.D)) + 256, dtype=torch.uint8, device="cuda")
            wp = (w.data_ptr() + 255) & ~255
            a.workspace, a.workspace_bytes = wp, w.numel() - (wp - w.data_ptr())      # each emulated rank keeps its own workspace
            own = bufs[r].data_ptr() + parity * KVH * L * 2
            a.head_scores = own + r * per * L * 2
            a.skip_select, a.skip_score, a.score_rows = 1, 0, KVH
            a.xchg_world, a.xchg_rank, a.xchg_epoch = world, r, epoch
            for s in range(world):
                a.xchg_scores[s] = bufs[s].data_ptr() + parity * KVH * L * 2
                a.xchg_flags[s] = flags[s].data_ptr() + parity * 8 * 4
            args.append(a)
            outs.append(o)
            ws.append((w, keepalive))
        for a in args:                                       # call 1 of every rank: score + put + flags
            _call(a)
        torch.cuda.synchronize()
        # what each emulated rank computes for its own heads, alone (the same launch shape: bit-identical by construction)
        solo = [lc.pivot_update(q[:, r * per * G:(r + 1) * per * G], k[:, r * per:(r + 1) * per], v[:, r * per:(r + 1) * per],
                                keep, mask, pos, rot, mrope, reforge)[4] for r in range(world)]
        rows = torch.cat(solo)
        for r in range(world):
            assert flags[r][parity, :world].tolist() == [epoch] * world
            got = bufs[r][parity].view(KVH, L)
            assert torch.equal(got, rows), "every rank holds every rank's rows after the exchange"
        # against the full-width launch the per-head rows may differ in the last bf16 bit: a unit split between CTAs folds
        # its fp32 partials in another order when the launch holds other heads
        d = (rows.view(torch.int16).int() - want_hs.view(torch.int16).int()).abs()
        assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 1e-3
        for r, a in enumerate(args):                         # call 2: wait, select on all rows, compact own heads
            a.skip_select, a.skip_score = 0, 1
            a.head_scores = bufs[r].data_ptr() + parity * KVH * L * 2
            _call(a)
        torch.cuda.synchronize()
        # the selection runs on the gathered rows: identical on every rank, and equal to the reference rule on those rows
        score = rows.float().mean(0).to(torch.bfloat16).masked_fill(mask, 1.0)
        ref_idx = torch.sort(score.float(), descending=True, stable=True).indices[:keep].sort().values
        for r, o in enumerate(outs):
            assert torch.equal(o["keep_idx"].long(), ref_idx)
            assert torch.equal(o["keep_idx"], outs[0]["keep_idx"])
        # against the single-GPU update: the same kept set, or one that differs only on the cut
        from helpers import index_parity
        full_score = want_hs.float().mean(0).to(torch.bfloat16).masked_fill(mask, 1.0)
        same, justified, _ = index_parity(outs[0]["keep_idx"], want_idx, full_score, keep)
        assert justified
        if same:
            for r, o in enumerate(outs):
                assert torch.equal(o["k_out"], want_k[:, r * per:(r + 1) * per])
                assert torch.equal(o["v_out"], want_v[:, r * per:(r + 1) * per])
                assert torch.equal(o["pos_out"], want_p)


def test_exchange_argument_errors():
    lc = _lc()
    from retake import _native as N
    q, k, v = qkv(4, 2, 64, 64, 1.0, seed=1)
    a, o, ka, _, _ = lc.fill_update_args(None, False, lc._inv_freq_on_device, q, k, v, None, None, None, 16)
    w = torch.empty(int(N.lib().rtk_pivot_update_workspace_bytes(4, 2, 64, 64)) + 256, dtype=torch.uint8, device="cuda")
    a.workspace, a.workspace_bytes = (w.data_ptr() + 255) & ~255, w.numel() - 256
    st = N.stream_ptr(torch.device("cuda"))
    a.skip_select = a.skip_score = 1
    assert N.lib().rtk_pivot_update(C.byref(a), st) == -1                      # both at once
    a.skip_score, a.xchg_world, a.xchg_rank = 0, 2, 0
    assert N.lib().rtk_pivot_update(C.byref(a), st) == -1                      # peers missing
    a.skip_select, a.xchg_world = 0, 2
    assert N.lib().rtk_pivot_update(C.byref(a), st) == -1                      # an exchange needs one of the two modes


@pytest.mark.parametrize("sync", [False, True])
def test_gather_owned_fills_exactly_the_owned_slots(sync):
    from retake import _native as N
    from retake import visual_compression as vc
    T, Np, Cc, t = 37, 48, 512, 15
    x = torch.randn(T, Np, Cc, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16).cuda()
    x[9] = x[8]
    want, mask, idx = vc.memory_bank_compress_keyframe(x[None], t, 3, sync=sync, return_indices=True)
    idx32 = idx.int().contiguous()
    total = torch.zeros_like(want)
    for t0, t1 in ((0, 11), (11, 12), (12, 37)):
        halo = int(t0 > 0)
        xl = x[t0 - halo:t1].contiguous()
        part = torch.zeros_like(want)
        N.check(N.lib().rtk_dpselect_gather_owned(xl.data_ptr(), xl.shape[0], t0 - halo, t0, t1, Np, Cc, idx32.data_ptr(), t,
                                                  int(sync), part.data_ptr(), N.stream_ptr(x.device)), "gather_owned")
        full = idx if idx.dim() == 2 else idx[:, None].expand(-1, Np)
        own = ((full >= t0) & (full < t1))[None, :, :, None].expand_as(want)
        assert torch.equal(part[own], want[own]) and not bool(part[~own].any())
        total += part
    assert torch.equal(total, want)
    # a window outside the frames held is refused
    assert N.lib().rtk_dpselect_gather_owned(x.data_ptr(), 5, 3, 2, 6, Np, Cc, idx32.data_ptr(), t, int(sync), total.data_ptr(),
                                             N.stream_ptr(x.device)) == -1
