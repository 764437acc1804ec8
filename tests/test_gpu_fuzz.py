"""Randomised shapes for the three selection kernels against torch's own CUDA ops (bit-exact): many (T, N, t) / (L, keep)
combinations with small value alphabets (heavy ties at the selection boundary), negative values, signed zeros, NaN."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _mods():
    from retake import longvideo_cache as lc
    from retake import visual_compression as vc
    return vc, lc


def _ref_dpselect_indices(dis, t, sync):
    """visual_compression.py:108-169 with torch-CUDA ops"""
    import torch.nn.functional as F
    T = dis.shape[0]
    rows = dis.mean(1)[None] if sync else dis.t().contiguous()
    arg = F.max_pool1d_with_indices(rows[:, None, :], 3, 1, padding=1)[1][:, 0]
    peak = arg == torch.arange(T, device=dis.device)[None]
    keys = torch.where(peak, rows + 2, rows)
    kept = torch.topk(keys, k=t, sorted=False, dim=1)[1].sort(dim=1)[0]
    if sync:
        idx = kept[0]
        return idx, peak[0][idx][:, None].repeat(1, dis.shape[1]).flatten()
    return kept.t(), peak.t().gather(0, kept.t()).flatten()


@pytest.mark.parametrize("seed", range(6))
def test_dpselect_select_fuzz(seed):
    vc, _ = _mods()
    g = torch.Generator().manual_seed(1000 + seed)
    for _ in range(8):
        T = int(torch.randint(1, 300, (1,), generator=g))
        N = int(torch.randint(1, 200, (1,), generator=g))
        t = int(torch.randint(1, T + 1, (1,), generator=g))
        levels = int(torch.randint(2, 40, (1,), generator=g))
        dis = (torch.randint(0, levels, (T, N), generator=g).float() / levels).cuda()
        dis[0] = 1.0
        if seed % 2:
            dis = dis - 0.25                                  # negative distances and both zeros
            dis[dis == 0] = -0.0
        for sync in (False, True):
            idx, mask = vc.dpselect_select(dis, t, sync)
            want_idx, want_mask = _ref_dpselect_indices(dis, t, sync)
            assert torch.equal(idx.long(), want_idx), (T, N, t, sync)
            assert torch.equal(mask, want_mask), (T, N, t, sync)


@pytest.mark.parametrize("seed", range(6))
def test_pivot_select_fuzz(seed):
    _, lc = _mods()
    g = torch.Generator().manual_seed(2000 + seed)
    for _ in range(10):
        L = int(torch.randint(1, 5000, (1,), generator=g))
        KVH = int(torch.randint(1, 9, (1,), generator=g))
        keep = int(torch.randint(1, L + 1, (1,), generator=g))
        levels = int(torch.randint(2, 50, (1,), generator=g))
        hs = (torch.randint(0, levels, (KVH, L), generator=g).float() / 32 - (0.5 if seed % 2 else 0.0)).to(BF).cuda()
        if seed == 5 and L > 4:
            hs[0, 3] = float("nan")                           # NaN sorts first in ATen's radix order
        mask = (torch.rand(L, generator=g) < 0.3).cuda() if seed % 3 else None
        idx, score = lc.pivot_select(hs, keep, mask, return_scores=True)
        s = hs.mean(0)
        assert torch.equal(score.view(torch.int16), s.view(torch.int16))
        if mask is not None:
            s = s.masked_fill(mask, 1.0)
        want = s.topk(keep).indices.sort().values
        assert torch.equal(idx.long(), want), (L, KVH, keep)


@pytest.mark.parametrize("seed", range(4))
def test_mallm_fuzz(seed):
    """random small banks with repeated frames (equal similarities): fused loop == the reference's op loop on the GPU"""
    from oracle import reference_ops as ro
    vc, _ = _mods()
    g = torch.Generator().manual_seed(3000 + seed)
    for _ in range(4):
        T = int(torch.randint(2, 40, (1,), generator=g))
        N = int(torch.randint(1, 300, (1,), generator=g))
        C = 256 + 8 * int(torch.randint(0, 64, (1,), generator=g))
        t = int(torch.randint(1, T + 1, (1,), generator=g))
        base = torch.randn(max(2, T // 3), N, C, generator=g)
        x = base[torch.randint(0, base.shape[0], (T,), generator=g)].to(BF)[None].cuda()       # many identical frames
        for sync in (False, True):
            for hard in (False, True):
                want, want_size = ro.mallm_compress(x.clone(), t, sync, hard)
                got, got_size = vc.mallm_compress(x, t, sync=sync, hard=hard)
                assert torch.equal(got, want), (T, N, C, t, sync, hard)
                assert hard or torch.equal(got_size, want_size), (T, N, C, t, sync, hard)


@pytest.mark.parametrize("seed", range(4))
def test_distance_fuzz_over_hidden_sizes(seed):
    """every template instantiation of the streaming cosine kernel (K4 buckets, exact-multiple and ragged rows), random C"""
    import torch.nn.functional as F
    vc, _ = _mods()
    g = torch.Generator().manual_seed(4000 + seed)
    fixed = [256, 512, 1152, 2048, 3584, 4096, 6144, 8192, 264, 1160, 3592, 7680, 8184]
    for i in range(8):
        C = fixed[(seed * 5 + i) % len(fixed)] if i < 5 else 256 + 8 * int(torch.randint(0, 993, (1,), generator=g))
        T = int(torch.randint(2, 24, (1,), generator=g))
        N = int(torch.randint(1, 40, (1,), generator=g))
        x = torch.randn(T, N, C, generator=g).to(BF)
        x[T // 2] = x[T // 2 - 1]                             # an exactly repeated frame: sim == 1, dis == 0
        x = x.cuda()
        got = vc.dpselect_distance(x)
        sim = F.cosine_similarity(x[:-1], x[1:], dim=-1)
        want = torch.cat([torch.ones_like(sim[:1], dtype=torch.float32), 1 - sim.float()], dim=0)
        assert torch.equal(got, want), (T, N, C)
        if T > 2:
            halo = vc.dpselect_distance(x[1:], halo=True)     # frames 2.. of the same video, frame 1 as halo
            assert torch.equal(halo, want[2:]), (T, N, C)


@pytest.mark.parametrize("seed", range(6))
def test_pivot_update_fuzz(seed):
    """random (H, KVH, L, D, ratio, reforge, mrope / 1-D, mask) through PivotKVCache.update: scores within 1 bf16 ulp of the
    torch-CUDA op sequence (O(1) logits; see the note below for huge ones), kept set consistent, kept V / K / positions exact
    (ragged L, L < 128, G in {1, 2, 4, 7})."""
    from helpers import TableRotary
    from oracle import pivotkv as op
    from test_gpu_pivotkv import _cfg, _check_keep, _lc, qkv, ref_head_scores_cuda, ulp_diff
    lc = _lc()
    g = torch.Generator().manual_seed(5000 + seed)
    for trial in range(5):
        D = (64, 128)[int(torch.randint(0, 2, (1,), generator=g))]
        KVH = (1, 2, 4, 8)[int(torch.randint(0, 4, (1,), generator=g))]
        G = (1, 2, 4, 6, 7, 8)[int(torch.randint(0, 6, (1,), generator=g))]
        H = KVH * G
        L = int(torch.randint(1, 2600, (1,), generator=g))
        ratio = float(torch.rand(1, generator=g)) * 0.9 + 0.05
        reforge = bool(torch.randint(0, 2, (1,), generator=g))
        use_mrope = bool(torch.randint(0, 2, (1,), generator=g))
        mrope = ([8, 12, 12] if D == 64 else [16, 24, 24]) if use_mrope else None
        rot = TableRotary(D, mrope=use_mrope)
        rot.inv_freq = rot.inv_freq.cuda()
        cache = lc.PivotKVCache(_cfg(H, KVH, D, 1, ratio, reforge))
        alpha = (1.0, 2.0)[trial % 2]
        q, k, v = qkv(H, KVH, L, D, alpha, seed=seed * 100 + trial)
        mask = (torch.rand(L, generator=g) < 0.25).cuda() if trial % 2 else None
        cache.kvcache_compression = True
        cache.keypatches_mask_chunk = mask
        ar = torch.arange(L)
        pos = (torch.stack([5 + ar // 64, (ar % 64) // 8, ar % 8])[:, None] if use_mrope else (5 + ar)[None]).cuda()
        tag = (H, KVH, L, D, round(ratio, 3), reforge, use_mrope)
        ko, vo = cache.update(k, v, 0, {"query_states": q, "position_ids": pos.clone(), "rotary_emb": rot, "mrope_section": mrope})
        keep = max(1, int(ratio * L))
        idx = cache.last_keep_indices.long()
        assert idx.numel() == keep and torch.equal(ko, k) and torch.equal(vo, v), tag
        qq, kk = q, k
        if reforge:
            cos, sin = rot(v, pos)
            c1 = op.select_mrope(cos, mrope) if use_mrope else cos
            s1 = op.select_mrope(sin, mrope) if use_mrope else sin
            qq = op.unrotate(q, c1, s1, rot.attention_scaling, "cuda")
            kk = op.unrotate(k, c1, s1, rot.attention_scaling, "cuda")
        ref_hs = ref_head_scores_cuda(qq, kk)
        hs = cache.last_head_scores
        if mask is not None:
            # key patches: their column sums are never computed (pass 2 skips them) - the exposed rows carry the 1.0 the
            # selection uses for them (longvideo_cache.py:272-274)
            assert bool((hs[:, mask] == 1.0).all()), tag
            d = ulp_diff(hs[:, ~mask], ref_hs[:, ~mask]) if bool((~mask).any()) else torch.zeros(1, dtype=torch.int32)
        else:
            d = ulp_diff(hs, ref_hs)
        if alpha == 1.0:
            assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 0.02, tag
        else:
            # logits of magnitude ~100 have a bf16 ulp of ~1: ONE logit that cuBLAS' fp32 accumulation order rounds the other
            # way moves a softmax weight by several per cent (tests/probes/probe_ulp2.py: the kernel then agrees with exactly
            # computed dot products) - rare entries may be 2-3 steps apart, the bulk stays within one
            assert int(d.max()) <= 4 and float((d > 1).float().mean()) < 1e-3 and float((d > 0).float().mean()) < 0.02, tag
        _check_keep(cache.last_keep_indices, cache.last_head_scores.mean(0), ref_hs.mean(0), mask, keep)
        assert torch.equal(cache.layers[0].values, v[:, :, idx]), tag
        if not reforge:
            assert torch.equal(cache.layers[0].keys, k[:, :, idx]), tag
        else:
            want_pos = pos[..., idx].clone()
            want_pos[0] = op.reforge_temporal(want_pos[0], keep, L)
            assert torch.equal(cache.position_cache[0], want_pos), tag
            cos, sin = rot(v, want_pos)
            c2 = op.select_mrope(cos, mrope) if use_mrope else cos
            s2 = op.select_mrope(sin, mrope) if use_mrope else sin
            assert torch.equal(cache.layers[0].keys, op.rotate(kk[:, :, idx], c2, s2)), tag


@pytest.mark.parametrize("deferred", [False, True])
@pytest.mark.parametrize("seed", range(4))
def test_cache_state_machine_fuzz(seed, deferred):
    """random interleavings of compressing / plain updates over two layers with ragged chunk lengths, long enough to grow
    the preallocated buffers several times: the cache must always equal [past kept rows | ...] built from the kernel's own
    kept indices, the returned tensors must equal [past | chunk] (deferred tail overwrite, buffer growth, key_cache views).
    ``deferred``: the compression itself is postponed to after_forward() / the layer's next update / the next read and runs
    batched over the layers that are pending at that moment."""
    from helpers import TableRotary
    from test_gpu_pivotkv import _cfg, _lc, qkv
    lc = _lc()
    g = torch.Generator().manual_seed(6000 + seed)
    H, KVH, D, layers = 4, 2, 64, 2
    ratio = (0.1, 0.3, 0.6, 0.9)[seed]
    rot = TableRotary(D, mrope=False)
    rot.inv_freq = rot.inv_freq.cuda()
    cfg = _cfg(H, KVH, D, layers, ratio, False)
    cfg.longvideo_kwargs["kvcache_compression_kwargs"]["deferred_compression"] = deferred
    cache = lc.PivotKVCache(cfg)
    want_k = [torch.empty(1, KVH, 0, D, dtype=BF, device="cuda") for _ in range(layers)]
    want_v = [torch.empty(1, KVH, 0, D, dtype=BF, device="cuda") for _ in range(layers)]
    evicted = [0] * layers
    pend = [None] * layers                 # (k, v, kept-index tensor) of a chunk whose kept rows are not folded into want_* yet

    def resolve(layer):
        if pend[layer] is not None:
            kk, vv, idx_t = pend[layer]
            idx = idx_t.long()
            assert idx.numel() == max(1, int(ratio * kk.shape[2])) and bool((idx[1:] > idx[:-1]).all())
            want_k[layer] = torch.cat([want_k[layer], kk[:, :, idx]], dim=2)
            want_v[layer] = torch.cat([want_v[layer], vv[:, :, idx]], dim=2)
            pend[layer] = None

    for step in range(14):
        compress = bool(torch.randint(0, 4, (1,), generator=g))            # 3 of 4 steps compress
        L = int(torch.randint(1, 3000, (1,), generator=g)) if compress else int(torch.randint(1, 40, (1,), generator=g))
        cache.kvcache_compression = compress
        cache.keypatches_mask_chunk = ((torch.rand(L, generator=g) < 0.2).cuda() if (compress and step % 2) else None)
        for layer in range(layers):
            q, k, v = qkv(H, KVH, L, D, 1.0, seed=seed * 1000 + step * 10 + layer)
            past_len = cache.get_seq_length(layer)
            pos = (past_len + torch.arange(L))[None].cuda()
            kw = {"position_ids": pos}
            if compress:
                kw.update({"query_states": q, "rotary_emb": rot, "mrope_section": None})
            ko, vo = cache.update(k, v, layer, kw)
            resolve(layer)                 # update() settles this layer's own debt before it appends
            assert want_k[layer].shape[2] == past_len
            assert torch.equal(ko, torch.cat([want_k[layer], k], dim=2)) and torch.equal(vo, torch.cat([want_v[layer], v], dim=2))
            if compress:
                if deferred:
                    assert cache._deferred and cache.layers[layer]._deferred_owner is cache
                    pend[layer] = (k, v, cache._deferred[-1]["outs"]["keep_idx"])
                else:
                    pend[layer] = (k, v, cache.last_keep_indices)
                    resolve(layer)
                evicted[layer] += L - max(1, int(ratio * L))
            else:
                want_k[layer] = torch.cat([want_k[layer], k], dim=2)
                want_v[layer] = torch.cat([want_v[layer], v], dim=2)
        if step % 3 == 2:
            cache.after_forward()
            assert not cache._deferred
            for layer in range(layers):
                resolve(layer)
        for layer in range(layers):
            keep_pending = 0 if pend[layer] is None else pend[layer][2].numel()
            assert cache.get_seq_length(layer) == want_k[layer].shape[2] + keep_pending      # lengths are settled at once
            if step % 2:
                got_k = cache.layers[layer].keys                                             # a read settles the bytes
                resolve(layer)
                assert torch.equal(got_k, want_k[layer]) and torch.equal(cache.key_cache[layer], want_k[layer])
                assert torch.equal(cache.layers[layer].values, want_v[layer])
    for layer in range(layers):
        got_k = cache.layers[layer].keys
        resolve(layer)
        assert torch.equal(got_k, want_k[layer]) and torch.equal(cache.layers[layer].values, want_v[layer])
        assert cache.num_evicted_tokens[layer] == evicted[layer]
