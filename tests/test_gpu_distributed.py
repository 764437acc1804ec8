"""2-GPU NCCL run of the frame-range-sharded DPSelect and the KV-head-sharded PivotKV update against the single-GPU
operators (bit-exact).  Needs >= 2 visible GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port):
    for p in (ROOT, os.path.join(ROOT, "video-retake_b200"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    from helpers import TableRotary, scene_video
    from retake import distributed as rd
    from retake import longvideo_cache as lc
    from retake import visual_compression as vc
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    try:
        g = torch.Generator().manual_seed(3)
        T, N, C = 65, 96, 1152
        x = scene_video(g, T, N, C, dup_every=6).to(torch.bfloat16).to(dev)
        t0, t1 = rd.split_range(T, world)[rank]
        for sync in (False, True):
            for t in (T, 32, 9):
                want_out, want_mask, want_idx = vc.memory_bank_compress_keyframe(x[None], t, 3, sync=sync, return_indices=True)
                rows, slots, mask, idx = rd.dpselect_frame_sharded(x[t0 - int(t0 > 0):t1], t0, t1, T, t, sync)
                assert torch.equal(mask, want_mask) and torch.equal(idx.long(), want_idx)
                out = rd.assemble_compacted(rows, slots, t, N)
                assert torch.equal(out, want_out)
                # sync-free form: one all_gather_into_tensor, owned rows written straight into their slots
                part, mask2, idx2 = rd.dpselect_frame_sharded_fused(x[t0 - int(t0 > 0):t1], t0, t1, T, t, sync, zero_fill=True)
                assert torch.equal(mask2, want_mask) and torch.equal(idx2.long(), want_idx)
                full = idx2.long() if idx2.dim() == 2 else idx2.long()[:, None].expand(-1, N)
                own = ((full >= t0) & (full < t1))[None, :, :, None].expand_as(want_out)
                assert torch.equal(part[own], want_out[own]) and not bool(part[~own].any())
                assert torch.equal(rd.assemble_owned(part), want_out)

        H, KVH, L, D, mrope = 28, 4, 1024, 128, [16, 24, 24]
        q = torch.randn(1, L, H, D, generator=g).to(torch.bfloat16).to(dev).transpose(1, 2)
        k = torch.randn(1, L, KVH, D, generator=g).to(torch.bfloat16).to(dev).transpose(1, 2)
        v = torch.randn(1, L, KVH, D, generator=g).to(torch.bfloat16).to(dev).transpose(1, 2)
        ar = torch.arange(L, device=dev)
        pos = torch.stack([7 + ar // 256, (ar % 256) // 16, ar % 16])[:, None]
        rot = TableRotary(D)
        rot.inv_freq = rot.inv_freq.to(dev)
        mask = (torch.rand(L, generator=g) < 0.2).to(dev)
        keep = 256
        per = [KVH // world] * world
        g0, G = rank * per[0], H // KVH
        for reforge in (False, True):
            import types
            cfg = types.SimpleNamespace(hidden_size=H * D, num_hidden_layers=1, num_attention_heads=H, num_key_value_heads=KVH)
            cfg.longvideo_kwargs = {"kvcache_compression": True, "kvcache_compression_kwargs": {
                "compression_ratio": keep / L, "compression_method": "pivotkv", "pos_embed_reforge": reforge}}
            cache = lc.PivotKVCache(cfg)
            cache.keypatches_mask_chunk = mask
            cache.update(k, v, 0, {"query_states": q, "position_ids": pos.clone(), "rotary_emb": rot, "mrope_section": mrope})
            for transport in ("p2p", "nccl", "unfused"):
                for rep in range(3):                    # several epochs: both buffer parities, flags re-used
                    if transport == "unfused":
                        fn = rd._pivot_update_kv_sharded_unfused
                        extra = {}
                    else:
                        fn, extra = rd.pivot_update_kv_sharded, {"transport": transport}
                    kk, vv, pp, idx, hs = fn(q[:, g0 * G:(g0 + per[0]) * G], k[:, g0:g0 + per[0]], v[:, g0:g0 + per[0]], keep,
                                             per, mask, pos.clone(), rot, mrope, reforge, **extra)
                    # per-head rows: last-bit differences against the full-width launch are possible (fp32 fold order of a
                    # unit split between CTAs); the kept set is the single-GPU one or differs on the cut
                    # (key-patch columns are left out: the fused paths write the reference's 1.0 there without computing
                    # them, the unfused rtk_pivot_score computes every column - section 4 "key elision")
                    d = (hs.view(torch.int16).int() - cache.last_head_scores.view(torch.int16).int()).abs()[:, ~mask]
                    assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 1e-3, (transport, rep)
                    from helpers import index_parity
                    full_score = cache.last_head_scores.float().mean(0).to(torch.bfloat16).masked_fill(mask, 1.0)
                    same, justified, _ = index_parity(idx, cache.last_keep_indices, full_score, keep)
                    assert justified, (transport, rep)
                    if same:
                        assert torch.equal(kk, cache.layers[0].keys[:, g0:g0 + per[0]])
                        assert torch.equal(vv, cache.layers[0].values[:, g0:g0 + per[0]])
                        if reforge:
                            assert torch.equal(pp, cache.position_cache[0])
            if rank == 0:
                ex = rd.ScoreExchange.get(None, dev, KVH, L, "p2p")
                print("score exchange transport:", ex.transport, flush=True)
            # all layers of a chunk at once: two batched calls around one exchange
            if reforge and lc._rotary_inv_freq(rot) is None:
                continue
            n_layers = 5
            data = []
            for layer in range(n_layers):
                gl = torch.Generator().manual_seed(100 + layer)
                ql = torch.randn(1, L, H, D, generator=gl).to(torch.bfloat16).to(dev).transpose(1, 2)
                kl = torch.randn(1, L, KVH, D, generator=gl).to(torch.bfloat16).to(dev).transpose(1, 2)
                vl = torch.randn(1, L, KVH, D, generator=gl).to(torch.bfloat16).to(dev).transpose(1, 2)
                ml = (torch.rand(L, generator=gl) < 0.2).to(dev) if layer % 2 == 0 else None
                data.append((ql, kl, vl, ml))
            want = [lc.pivot_update(ql, kl, vl, keep, ml, pos.clone(), rot, mrope, reforge) for ql, kl, vl, ml in data]
            for transport in ("p2p", "nccl"):
                for rep in range(2):
                    got = rd.pivot_update_batch_kv_sharded(
                        [(ql[:, g0 * G:(g0 + per[0]) * G], kl[:, g0:g0 + per[0]], vl[:, g0:g0 + per[0]], ml, pos.clone())
                         for ql, kl, vl, ml in data], keep, per, rot, mrope, reforge, transport=transport)
                    for (kk, vv, pp, idx), (wk, wv, wp, widx, whs), (_, _, _, ml) in zip(got, want, data):
                        full_score = whs.float().mean(0).to(torch.bfloat16)
                        if ml is not None:
                            full_score = full_score.masked_fill(ml, 1.0)
                        same, justified, _ = index_parity(idx, widx, full_score, keep)
                        assert justified, (transport, rep)
                        if same:
                            assert torch.equal(kk, wk[:, g0:g0 + per[0]]) and torch.equal(vv, wv[:, g0:g0 + per[0]])
                            assert torch.equal(pp, wp)
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharded_paths_match_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)
