"""GPU: pass 2 of the scoring skips the key patches (their score is overwritten with 1.0 before the top-k,
``longvideo_cache.py:272-274``, so their column sums are never read).  The kept indices, K / V rows and positions must be what
the reference's op sequence gives with the mask, for every mask density incl. the extremes, ragged lengths, both update
paths (immediate and batched) and with re-forging."""
import pytest
import torch

from helpers import TableRotary, index_parity
from oracle import reference_ops as ro
from test_gpu_pivotkv import _cfg, qkv, ref_head_scores_cuda, ulp_diff

pytestmark = pytest.mark.gpu


def _lc():
    from retake import longvideo_cache as lc
    return lc


def _mask(L, kind, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "none":
        return torch.zeros(L, dtype=torch.bool)
    if kind == "all":
        return torch.ones(L, dtype=torch.bool)
    if kind == "all_but_one":
        m = torch.ones(L, dtype=torch.bool)
        m[L // 2] = False
        return m
    if kind == "block":                                   # a whole 128-key tile masked, the rest free
        m = torch.zeros(L, dtype=torch.bool)
        m[: min(L, 128)] = True
        return m
    return torch.rand(L, generator=g) < float(kind)


@pytest.mark.parametrize("kind", ["none", "all", "all_but_one", "block", "0.3", "0.9"])
@pytest.mark.parametrize("H,KVH,L,D", [(28, 4, 1000, 128), (4, 2, 130, 64), (8, 8, 257, 128)])
def test_masked_keys_are_skipped_not_mis_scored(H, KVH, L, D, kind):
    lc = _lc()
    mask = _mask(L, kind, 7).cuda()
    q, k, v = qkv(H, KVH, L, D, 1.0, seed=L + H)
    keep = max(1, L // 3)
    kk, vv, _, idx, hs = lc.pivot_update(q, k, v, keep, mask, None, None, None, False)
    ref = ref_head_scores_cuda(q, k)
    assert bool((hs[:, mask] == 1.0).all())
    if bool((~mask).any()):
        d = ulp_diff(hs[:, ~mask], ref[:, ~mask])
        assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 0.02
    _, _, _, idx_ref, score_ref = ro.pivot_update(q, k, v, keep / L + 1e-9, mask, None, None, None, False)
    assert idx_ref.numel() == keep
    same, justified, _ = index_parity(idx, idx_ref, score_ref, keep)
    assert justified
    assert torch.equal(kk, k[:, :, idx.long()]) and torch.equal(vv, v[:, :, idx.long()])


def test_batched_layers_with_different_masks():
    """rtk_pivot_update_batch: every layer brings its own mask (or none); each layer's pass 2 visits its own key count"""
    lc = _lc()
    H, KVH, L, D, layers, ratio = 28, 4, 700, 128, 5, 0.25
    rot = TableRotary(D)
    rot.inv_freq = rot.inv_freq.cuda()
    mrope = [16, 24, 24]
    ar = torch.arange(L, device="cuda")
    pos = torch.stack([3 + ar // 64, (ar % 64) // 8, ar % 8])[:, None]
    kinds = ["0.3", "none", "all", "0.9", "block"]
    outs = {}
    for deferred in (False, True):
        cfg = _cfg(H, KVH, D, layers, ratio, True)
        cfg.longvideo_kwargs["kvcache_compression_kwargs"]["deferred_compression"] = deferred
        cache = lc.PivotKVCache(cfg)
        cache.kvcache_compression = True
        res = []
        for layer in range(layers):
            q, k, v = qkv(H, KVH, L, D, 3.0, seed=50 + layer)
            m = _mask(L, kinds[layer], layer).cuda()
            cache.keypatches_mask_chunk = None if kinds[layer] == "none" else m
            cache.update(k, v, layer, {"query_states": q, "position_ids": pos.clone(), "rotary_emb": rot, "mrope_section": mrope})
            if not deferred:
                res.append((cache.last_keep_indices.clone(), cache.last_head_scores.clone()))
        cache.after_forward()
        outs[deferred] = (cache, res)
    a, b = outs[False][0], outs[True][0]
    for layer in range(layers):
        assert torch.equal(a.layers[layer].keys, b.layers[layer].keys)
        assert torch.equal(a.layers[layer].values, b.layers[layer].values)
        assert torch.equal(a.position_cache[layer], b.position_cache[layer])
    # and against the reference's op sequence, layer by layer
    for layer in range(layers):
        q, k, v = qkv(H, KVH, L, D, 3.0, seed=50 + layer)
        m = None if kinds[layer] == "none" else _mask(L, kinds[layer], layer).cuda()
        _, _, _, idx_ref, score_ref = ro.pivot_update(q, k, v, ratio, m, pos.clone(), rot, mrope, True)
        idx = outs[False][1][layer][0]
        same, justified, _ = index_parity(idx, idx_ref, score_ref, int(ratio * L))
        assert justified, layer
