"""GPU probe: touch every kernel once at small, ragged sizes (run under compute-sanitizer memcheck / racecheck)."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "video-retake_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from helpers import TableRotary, scene_video
from retake import longvideo_cache as lc
from retake import visual_compression as vc

g = torch.Generator().manual_seed(0)
for (T, N, C) in ((9, 5, 256), (7, 33, 1152), (5, 3, 3584)):
    x = scene_video(g, T, N, C, dup_every=3).to(torch.bfloat16).cuda()
    for sync in (False, True):
        for t in (T, max(1, T // 2), 1):
            out, mask, idx = vc.memory_bank_compress_keyframe(x[None], t, 3, sync=sync, return_indices=True)
    d = vc.dpselect_distance(x[1:], halo=True)
    for sync in (False, True):                                   # MA-LLM compressors, soft and hard
        for hard in (False, True):
            for t in (T - 1, max(1, T // 2), 1):
                vc.mallm_compress(x[None], t, sync=sync, hard=hard)
xm = scene_video(g, 11, 137, 256, dup_every=3).to(torch.bfloat16).cuda()      # N % 8 != 0, N >= 128: position-dependent mean
for hard in (False, True):
    vc.mallm_compress(xm[None], 4, sync=True, hard=hard)
for (H, KVH, L, D) in ((4, 2, 130, 64), (28, 4, 200, 128), (8, 8, 1, 128)):
    cfg = types.SimpleNamespace(hidden_size=H * D, num_hidden_layers=2, num_attention_heads=H, num_key_value_heads=KVH)
    for reforge, deferred in ((False, False), (True, False), (False, True), (True, True)):
        cfg.longvideo_kwargs = {"kvcache_compression": True, "kvcache_compression_kwargs": {
            "compression_ratio": 0.3, "compression_method": "pivotkv", "pos_embed_reforge": reforge,
            "deferred_compression": deferred}}                  # deferred: the batched kernels (rtk_pivot_update_batch)
        cache = lc.PivotKVCache(cfg)
        rot = TableRotary(D)
        rot.inv_freq = rot.inv_freq.cuda()
        sec = [D // 8, 3 * D // 16, 3 * D // 16]
        for chunk in range(2):
            for layer in range(2):
                q = torch.randn(1, L, H, D, generator=g).to(torch.bfloat16).cuda().transpose(1, 2)
                k = torch.randn(1, L, KVH, D, generator=g).to(torch.bfloat16).cuda().transpose(1, 2)
                v = torch.randn(1, L, KVH, D, generator=g).to(torch.bfloat16).cuda().transpose(1, 2)
                ar = torch.arange(L, device="cuda")
                pos = torch.stack([chunk * 10 + ar // 16, (ar % 16) // 4, ar % 4])[:, None]
                cache.keypatches_mask_chunk = (torch.rand(L, generator=g) < 0.3).cuda()
                cache.update(k, v, layer, {"query_states": q, "position_ids": pos, "rotary_emb": rot, "mrope_section": sec})
            cache.after_forward()
        _ = cache.layers[0].keys.sum().item()
# round 2: the two-call KV-head split with both ranks emulated on this GPU (put kernel, flags, waiting select), the owned-slot
# compaction of the frame-range split and the one-call DPSelect operator - the same code as tests/test_gpu_exchange.py
import test_gpu_exchange as tx
for L_, reforge_ in ((333, True), (130, False)):
    tx.test_two_emulated_ranks_equal_single_gpu(L_, reforge_)
for sync_ in (False, True):
    tx.test_gather_owned_fills_exactly_the_owned_slots(sync_)
torch.cuda.synchronize()
print("sanitize pass done")
