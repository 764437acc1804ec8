"""CPU: host logic of the monkeypatch layer (no kernels): config patching, segmentation, chunk sizes, M-RoPE ids,
sequence truncation bookkeeping of compress_video_tokens with visual compression off."""
import types

import pytest
import torch

from retake import monkeypatch


def test_patch_config_functions_attach_longvideo_kwargs_and_yarn():
    exp = {"method": "retake", "scaling_factor": 4, "longvideo_kwargs": {"chunked_prefill_frames": 32}}
    cfg = types.SimpleNamespace(rope_scaling={"type": "mrope", "mrope_section": [16, 24, 24]})
    out = monkeypatch.patch_qwen2vl_config(cfg, exp)
    assert out is cfg and cfg.longvideo_kwargs == {"chunked_prefill_frames": 32}
    assert cfg.rope_scaling == {"mrope_section": [16, 24, 24], "rope_type": "yarn", "factor": 4, "beta_fast": 32.0, "beta_slow": 1.0}
    cfg2 = types.SimpleNamespace(text_config=types.SimpleNamespace(rope_parameters={"rope_type": "default", "rope_theta": 1e6}))
    monkeypatch.patch_llava_onevision_config(cfg2, exp)
    assert cfg2.text_config.rope_parameters == {"rope_type": "yarn", "factor": 4, "beta_fast": 32.0, "beta_slow": 1.0, "rope_theta": 1e6}
    cfg3 = types.SimpleNamespace(rope_scaling={"type": "mrope"})
    monkeypatch.patch_qwen2vl_config(cfg3, {})
    assert cfg3.longvideo_kwargs == {} and cfg3.rope_scaling == {"type": "mrope"}


def test_unknown_method_raises_like_the_reference():
    with pytest.raises(NotImplementedError):
        monkeypatch.patch_qwen2vl("h2o")
    with pytest.raises(NotImplementedError):
        monkeypatch.patch_llava_onevision("h2o")


def test_qwen2vl_helpers():
    from retake import qwen2_vl as g
    me = types.SimpleNamespace(config=types.SimpleNamespace(video_token_id=7, longvideo_kwargs={"chunked_prefill_frames": 32},
                                                            vision_config=types.SimpleNamespace(spatial_merge_size=2, temporal_patch_size=2)))
    ids = torch.tensor([[1, 2, 7, 7, 7, 7, 3, 7, 7, 4]])
    assert g.retake_Qwen2VLForConditionalGeneration_segment_input_ids(me, ids) == [
        (0, 2, "text"), (2, 6, "video"), (6, 7, "text"), (7, 9, "video"), (9, 10, "text")]
    thw = torch.tensor([[1024, 32, 32]])
    assert g.retake_Qwen2VLForConditionalGeneration_get_chunk_size(me, me.config, thw) == 4096      # 2048 frames @448px
    assert g.retake_Qwen2VLForConditionalGeneration_get_chunk_size(me, me.config, torch.tensor([[4, 32, 32]])) == 512
    me.config.longvideo_kwargs = {}
    assert g.retake_Qwen2VLForConditionalGeneration_get_chunk_size(me, me.config, thw) is None
    # M-RoPE ids: 2 text, video T=2 H=2 W=2 (merged), 3 text
    ids = torch.tensor([[1, 2] + [7] * 8 + [3, 4, 5]])
    pos, delta = g.mrope_position_ids(ids, 7, torch.tensor([[2, 4, 4]]), 2)
    assert pos[0, 0].tolist() == [0, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 5, 6]
    assert pos[1, 0].tolist() == [0, 1, 2, 2, 3, 3, 2, 2, 3, 3, 4, 5, 6]
    assert pos[2, 0].tolist() == [0, 1, 2, 3, 2, 3, 2, 3, 2, 3, 4, 5, 6]
    assert int(delta) == 7 - 13
    # visual compression off: everything passes through, no mask
    me.config.longvideo_kwargs = {"visual_compression": False}
    out = g.retake_Qwen2VLForConditionalGeneration_compress_video_tokens(me, input_ids=ids, attention_mask=None,
                                                                         video_embeds=torch.zeros(8, 4), position_ids=pos,
                                                                         video_grid_thw=torch.tensor([[2, 4, 4]]))
    assert out[0] is ids and out[4] is pos and out[6] is None


def test_llava_helpers():
    from retake import llava_onevision as g
    me = types.SimpleNamespace(config=types.SimpleNamespace(video_token_index=9, longvideo_kwargs={"chunked_prefill_frames": 32},
                                                            vision_config=types.SimpleNamespace(patch_size=14, image_size=384)))
    px = torch.zeros(1, 64, 3, 384, 384, dtype=torch.bfloat16)
    assert g.retake_LlavaOnevisionForConditionalGeneration_get_chunk_size(me, me.config, px) == 32 * 14 * 14   # 6272
    ids = torch.tensor([[9, 9, 1, 2]])
    assert g.retake_LlavaOnevisionForConditionalGeneration_segment_input_ids(me, ids) == [(0, 2, "video"), (2, 4, "text")]


# ---------------------------------------------------------------------------------------------------------------------
# Work partition of the scoring kernels (csrc/pivot_score.cu, struct TileRange) restated in Python: the invariants the
# kernel's partial-result protocol relies on, over the shapes the tests and the benchmark use.
def _tile_ranges(heads_per_layer, nt, layers, grid, pair, nt_a_per_layer=None):
    """[(cta, layer, unit-of-layer, tb0, tb1)] in the order a CTA walks them (TileRange::next); ``nt_a_per_layer``: stationary
    tiles per layer when pass 2 elides the key patches (default: nt, every key)"""
    out, meta = [], []
    for layer in range(layers):
        nt_a = nt if nt_a_per_layer is None else nt_a_per_layer[layer]
        nta = (nt_a + pair - 1) // pair
        Gl = heads_per_layer * nta * nt
        geff = max(1, min(grid, Gl // ((nt + 1) // 2)))
        meta.append((nta, Gl, geff))
    for cta in range(grid):
        for layer in range(layers):
            nta, Gl, geff = meta[layer]
            if cta >= geff:
                continue
            g, g1 = Gl * cta // geff, Gl * (cta + 1) // geff
            while g < g1:
                ul = g // nt
                tb0 = g - ul * nt
                tb1 = min(nt, tb0 + (g1 - g))
                out.append((cta, layer, ul, tb0, tb1))
                g += tb1 - tb0
    return out, meta


def test_score_kernel_partition_invariants():
    import itertools
    shapes = [(28, 32, 1, None), (28, 32, 28, None), (28, 49, 3, None), (4, 8, 2, None), (4, 1, 5, None), (14, 1, 1, None),
              (8, 2, 33, None), (28, 18, 2, None), (2, 3, 1, None), (7, 32, 28, None), (14, 32, 2, None), (7, 49, 1, None),
              (1, 5, 1, None),
              # pass 2 with key elision: fewer stationary tiles than streamed ones, different per layer, down to none
              (28, 32, 3, [23, 22, 24]), (7, 32, 4, [23, 1, 0, 32]), (4, 8, 3, [5, 8, 1]), (28, 49, 2, [35, 2]), (2, 3, 2, [1, 0])]
    for (H, nt, layers, nt_a), pair in itertools.product(shapes, (1, 2)):
        nta_full = (nt + pair - 1) // pair
        # host side (score_launch): one layer's units AT THE FULL KEY COUNT decide the grid; a range is at least half a unit
        # long, so launches with fewer units than SMs (7 heads per rank of the KV-head split) still use the whole machine
        grid = max(1, min(H * nta_full * nt // ((nt + 1) // 2), 148))
        steps, meta = _tile_ranges(H, nt, layers, grid, pair, nt_a)
        seen = {}
        for cta, layer, ul, tb0, tb1 in steps:
            nta, Gl, geff = meta[layer]
            assert 0 <= tb0 < tb1 <= nt and 0 <= ul < H * nta
            seen.setdefault((layer, ul), []).append((cta, tb0, tb1))
        want_units = [(layer, ul) for layer in range(layers) for ul in range(H * meta[layer][0])]
        assert sorted(seen) == want_units, "every unit of every layer is visited (and nothing else)"
        for (layer, ul), parts in seen.items():
            nta, Gl, geff = meta[layer]
            parts.sort(key=lambda p: p[1])
            # the streamed tiles of a unit are covered exactly once, by at most THREE CTAs: the piece that starts at tile 0
            # writes partial 0, a piece that neither starts nor ends the unit partial 1, the piece that ends it partial 2; the
            # first piece clears the partials nobody writes (kernel: fill1 / fill2)
            assert len(parts) <= 3 and parts[0][1] == 0 and parts[-1][2] == nt
            assert all(a[2] == b[1] for a, b in zip(parts, parts[1:]))
            assert len({p[0] for p in parts}) == len(parts)
            written = set()
            for cta, tb0, tb1 in parts:
                first, last = tb0 == 0, tb1 == nt
                part = 0 if first else (2 if last else 1)
                assert part not in written
                written.add(part)
                if first:
                    next_reaches_end = Gl * (cta + 2) // geff >= (ul + 1) * nt      # TileRange::next_cta_reaches_end_of
                    if last or next_reaches_end:
                        assert 1 not in written
                        written.add(1)
                    if last:
                        written.add(2)
            assert written == {0, 1, 2}, "every partial plane of the unit is written exactly once"
        # batched launches cut every layer where a single-layer launch (of that layer's key count) cuts it
        for layer in range(layers):
            one, _ = _tile_ranges(H, nt, 1, grid, pair, None if nt_a is None else [nt_a[layer]])
            mine = [(c, u, a, b) for c, l, u, a, b in steps if l == layer]
            assert mine == [(c, u, a, b) for c, _, u, a, b in one]
        # every CTA that takes part in a layer gets the same number of tile-steps, +-1
        for layer in range(layers):
            per = {}
            for cta, l, u, tb0, tb1 in steps:
                if l == layer:
                    per[cta] = per.get(cta, 0) + tb1 - tb0
            if per:
                assert len(per) == min(meta[layer][2], sum(1 for _ in per)) and max(per.values()) - min(per.values()) <= 1


def test_rebased_position_ids_equal_the_per_layer_rule():
    """PivotKVCache.rebased_position_ids: the ids of a chunk for all layers at once == the attention forward's per-layer
    `ids[0] += prev_l + 1 - ids[0, ..., 0]` (reference qwen2_vl.py:68-73, llava_onevision.py:80-89)"""
    import types
    from retake.longvideo_cache import PivotKVCache
    cfg = types.SimpleNamespace(hidden_size=64, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2)
    cfg.longvideo_kwargs = {"kvcache_compression": True, "kvcache_compression_kwargs": {
        "compression_ratio": 0.5, "compression_method": "pivotkv", "pos_embed_reforge": True}}
    for mrope in (True, False):
        cache = PivotKVCache(cfg)
        L = 10
        ar = torch.arange(L)
        pos = torch.stack([40 + ar // 4, ar % 4, ar % 2])[:, None] if mrope else (40 + ar)[None]
        # empty cache: every layer starts at temporal id 0
        got = cache.rebased_position_ids(pos, 3)
        for layer in range(3):
            want = pos.clone()
            want[0] += -1 + 1 - pos[0].reshape(-1)[0]
            assert torch.equal(got[layer], want)
        # layers whose caches end at different temporal ids
        for layer, last in enumerate((7, 3, 11)):
            if mrope:
                kept = torch.stack([torch.tensor([0, 2, last]), torch.tensor([0, 1, 1]), torch.tensor([0, 0, 1])])[:, None]
            else:
                kept = torch.tensor([[0, 2, last]])
            cache.update_position_ids(kept, layer)
        got = cache.rebased_position_ids(pos, 3)
        assert got.shape == (3,) + tuple(pos.shape)
        for layer, last in enumerate((7, 3, 11)):
            want = pos.clone()
            want[0] += last + 1 - pos[0].reshape(-1)[0]
            assert torch.equal(got[layer], want)
            assert int(got[layer][0].reshape(-1)[0]) == last + 1
        got[0][0] += 1000                                            # a layer's slice is its own memory
        assert int(got[1][0].reshape(-1)[0]) == 3 + 1
