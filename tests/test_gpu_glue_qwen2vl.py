"""GPU: the transformers-5 Qwen2-VL glue (retake/qwen2_vl.py) on a tiny random-init model (BASELINE config 1 shape:
2 layers, hidden 256, 2 KV heads, head dim 64): chunked prefill without compression equals one-shot prefill through
the stock language model; with DPSelect + PivotKV the cache shrinks as configured and generate() keeps working."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
VIDEO_ID, VSTART, VEND = 900, 901, 902


def tiny_model():
    from transformers import Qwen2VLConfig, Qwen2VLForConditionalGeneration
    cfg = Qwen2VLConfig(
        text_config=dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                         num_key_value_heads=2, vocab_size=1000, max_position_embeddings=4096,
                         rope_parameters={"rope_type": "default", "rope_theta": 10000.0, "mrope_section": [8, 12, 12]}),
        vision_config=dict(depth=2, embed_dim=64, hidden_size=256, num_heads=4, mlp_ratio=2, patch_size=14,
                           spatial_merge_size=2, temporal_patch_size=2, in_channels=3),
        video_token_id=VIDEO_ID, vision_start_token_id=VSTART, vision_end_token_id=VEND, image_token_id=903)
    cfg._attn_implementation = "sdpa"
    torch.manual_seed(0)
    model = Qwen2VLForConditionalGeneration(cfg).to(torch.bfloat16).cuda().eval()
    return model


def make_inputs(T=16, H=8, W=8, pre=5, post=6):
    g = torch.Generator().manual_seed(1)
    n_vid = T * (H // 2) * (W // 2)
    ids = torch.cat([torch.randint(0, 800, (pre,), generator=g), torch.tensor([VSTART]),
                     torch.full((n_vid,), VIDEO_ID), torch.tensor([VEND]), torch.randint(0, 800, (post,), generator=g)])
    # scene-structured frames so that DPSelect has something to find
    scene = torch.randn(1, H * W, 1176, generator=g)
    px = (scene.repeat(T, 1, 1) + 0.3 * torch.randn(T, H * W, 1176, generator=g))
    px[T // 2:] += torch.randn(1, H * W, 1176, generator=g)
    return dict(input_ids=ids[None].cuda(), pixel_values_videos=px.reshape(-1, 1176).to(torch.bfloat16).cuda(),
                video_grid_thw=torch.tensor([[T, H, W]]).cuda(), attention_mask=torch.ones(1, ids.numel(), dtype=torch.long).cuda())


def lv_kwargs(rv=1.0, rkv=1.0, reforge=False, chunk_frames=8, kv=True, deferred=False):
    return {"frame_chunk_size": 8, "chunked_prefill_frames": chunk_frames, "visual_compression": True,
            "visual_compression_kwargs": {"compression_ratio": rv, "compression_method": "Keyframe", "patch_sync": False,
                                          "return_keyframe_mask": True},
            "kvcache_compression": kv,
            "kvcache_compression_kwargs": {"dynamic_compression_ratio": False, "compression_ratio": rkv,
                                           "compression_method": "pivotkv", "pos_embed_reforge": reforge,
                                           "deferred_compression": deferred}}


@pytest.fixture()
def patched():
    from retake import monkeypatch, qwen2_vl
    monkeypatch.patch_qwen2vl("retake")
    yield qwen2_vl
    qwen2_vl.uninstall()


def test_helpers_and_config_patch(patched):
    from retake import monkeypatch
    model = tiny_model()
    cfg = monkeypatch.patch_qwen2vl_config(model.config, {"scaling_factor": 4, "longvideo_kwargs": lv_kwargs()})
    assert cfg.longvideo_kwargs["chunked_prefill_frames"] == 8
    assert cfg.text_config.rope_parameters["rope_type"] == "yarn" and cfg.text_config.rope_parameters["factor"] == 4
    inp = make_inputs()
    m = model.model
    assert m.get_chunk_size(model.config, inp["video_grid_thw"]) == 8 * 8 * 8 // 8
    segs = m.segment_input_ids(inp["input_ids"])
    assert [k for _, _, k in segs] == ["text", "video", "text"] and segs[1] == (6, 6 + 256, "video")
    pos, delta = patched.mrope_position_ids(inp["input_ids"], VIDEO_ID, inp["video_grid_thw"], 2)
    assert pos.shape == (3, 1, inp["input_ids"].shape[1])
    assert pos[0, 0, 6:262].tolist() == [6 + i // 16 for i in range(256)]                 # temporal: +1 per grid
    assert pos[1, 0, 6:22].tolist() == [6 + (i // 4) for i in range(16)] and pos[2, 0, 6:10].tolist() == [6, 7, 8, 9]
    assert int(pos[0, 0, 262]) == 6 + 16 and int(delta) == int(pos.max()) + 1 - inp["input_ids"].shape[1]
    with pytest.raises(NotImplementedError):
        monkeypatch.patch_qwen2vl("other")


def test_chunked_prefill_without_compression_equals_one_shot(patched):
    model = tiny_model()
    inp = make_inputs()
    outs = {}
    for name, chunk_frames in (("chunked", 8), ("one_shot", 1000)):
        model.config.longvideo_kwargs = lv_kwargs(1.0, 1.0, False, chunk_frames, kv=False)
        with torch.no_grad():
            o = model(**inp, use_cache=True)
        outs[name] = o
        assert o.past_key_values.get_seq_length() == inp["input_ids"].shape[1]
    a, b = outs["chunked"].logits[0, -1].float(), outs["one_shot"].logits[0, -1].float()
    assert torch.allclose(a, b, atol=0.08, rtol=0.05), float((a - b).abs().max())
    # and equals the stock language model on the same embeddings / positions (no ReTaKe code in the loop)
    pos, _ = patched.mrope_position_ids(inp["input_ids"], VIDEO_ID, inp["video_grid_thw"], 2)
    m = model.model
    with torch.no_grad():
        vid = m.visual(inp["pixel_values_videos"], grid_thw=inp["video_grid_thw"]).pooler_output
        emb = m.get_input_embeddings()(inp["input_ids"])
        vm = (inp["input_ids"] == VIDEO_ID).unsqueeze(-1).expand_as(emb)
        emb = emb.masked_scatter(vm, vid.to(emb.dtype))
        patched.uninstall()
        ref = m.language_model(inputs_embeds=emb, position_ids=pos, use_cache=True).last_hidden_state
        ref_logits = model.lm_head(ref[:, -1]).float()[0]
        patched.install()
    assert torch.allclose(a, ref_logits, atol=0.08, rtol=0.05), float((a - ref_logits).abs().max())


def test_deferred_compression_gives_the_same_prefill(patched):
    """deferred_compression: the chunk loop's after_forward() runs ONE batched compression per chunk; logits and caches
    equal compression inside update() bit for bit"""
    model = tiny_model()
    inp = make_inputs()
    outs = []
    for deferred in (False, True):
        model.config.longvideo_kwargs = lv_kwargs(rv=0.5, rkv=0.5, reforge=True, chunk_frames=8, deferred=deferred)
        with torch.no_grad():
            outs.append(model(**inp, use_cache=True))
    a, b = outs
    assert b.past_key_values.deferred_compression and not a.past_key_values.deferred_compression
    assert torch.equal(a.logits, b.logits)
    for l in range(2):
        assert torch.equal(a.past_key_values.layers[l].keys, b.past_key_values.layers[l].keys)
        assert torch.equal(a.past_key_values.layers[l].values, b.past_key_values.layers[l].values)
        assert torch.equal(a.past_key_values.position_cache[l], b.past_key_values.position_cache[l])


@pytest.mark.parametrize("deferred", [False, True])
@pytest.mark.parametrize("reforge", [False, True])
def test_compressed_prefill_and_generate(patched, reforge, deferred):
    from retake.longvideo_cache import PivotKVCache
    model = tiny_model()
    inp = make_inputs()
    model.config.longvideo_kwargs = lv_kwargs(rv=0.5, rkv=0.5, reforge=reforge, chunk_frames=8, deferred=deferred)
    with torch.no_grad():
        o = model(**inp, use_cache=True)
    cache = o.past_key_values
    assert isinstance(cache, PivotKVCache)
    # DPSelect keeps 8 of 16 grids -> 128 video tokens in two 64-token chunks; PivotKV keeps 32 of each
    want_len = 6 + 2 * 32 + 7
    assert [cache.get_seq_length(l) for l in range(2)] == [want_len, want_len]
    assert cache.num_evicted_tokens == [64, 64]
    assert cache.layers[0].keys.shape == (1, 2, want_len, 64) and cache.kvcache_compression is False
    if reforge:
        assert cache.position_cache[0].shape == (3, 1, want_len)
        t = cache.position_cache[0][0, 0]
        assert bool((t[1:] >= t[:-1]).all()), "temporal ids stay monotone after re-forging"
    assert o.logits.shape[1] == 7 and torch.isfinite(o.logits.float()).all()
    gen_cfg = copy.deepcopy(model.generation_config)
    gen_cfg.do_sample = False
    with torch.no_grad():
        out_ids = model.generate(**inp, max_new_tokens=4, generation_config=gen_cfg)
    assert out_ids.shape[1] == inp["input_ids"].shape[1] + 4


@pytest.mark.parametrize("method", ["MA-LLM", "MA-LLM-hard"])
def test_mallm_visual_compression_methods(patched, method):
    """`compression_method: MA-LLM / MA-LLM-hard` (qwen2_vl.py:402-409): the fused loop feeds the same chunked prefill."""
    from oracle import reference_ops as ro
    model = tiny_model()
    inp = make_inputs()
    kw = lv_kwargs(rv=0.5, rkv=0.5, chunk_frames=8)
    kw["visual_compression_kwargs"]["compression_method"] = method
    model.config.longvideo_kwargs = kw
    emb = torch.randn(16 * 16, model.config.text_config.hidden_size if hasattr(model.config, "text_config") else model.config.hidden_size,
                      generator=torch.Generator().manual_seed(5)).to(torch.bfloat16).cuda()
    ids, _, vid, _, _, _, mask = model.compress_video_tokens(input_ids=inp["input_ids"], attention_mask=inp["attention_mask"],
                                                             video_embeds=emb, video_grid_thw=torch.tensor([[16, 8, 8]]).cuda())
    want, _ = ro.mallm_compress(emb.reshape(1, 16, 16, -1).clone(), 8, False, method == "MA-LLM-hard")
    assert mask is None and torch.equal(vid, want.flatten(1, 2)[0])
    assert ids.shape[1] == inp["input_ids"].shape[1] - 8 * 16
    with torch.no_grad():
        o = model(**inp, use_cache=True)
    want_len = 6 + 2 * 32 + 7
    assert [o.past_key_values.get_seq_length(l) for l in range(2)] == [want_len, want_len]
    assert torch.isfinite(o.logits.float()).all()


def test_decode_positions_follow_the_original_prompt_with_visual_ratio_below_one(patched):
    """with DPSelect ratio < 1 the first decoded token must sit right after the largest prompt position
    (`cache_position[0] + rope_deltas`, both counted on the ORIGINAL prompt - reference qwen2_vl.py:583-589), not
    `num_token_diff` positions earlier"""
    model = tiny_model()
    inp = make_inputs()
    model.config.longvideo_kwargs = lv_kwargs(rv=0.5, rkv=0.5, reforge=True, chunk_frames=8)
    seen = []
    orig_lm = patched._lm

    def spy(self, cache, inputs_embeds, position_ids, **kw):
        seen.append(position_ids.clone())
        return orig_lm(self, cache, inputs_embeds, position_ids, **kw)

    patched._lm = spy
    try:
        gen_cfg = copy.deepcopy(model.generation_config)
        gen_cfg.do_sample = False
        with torch.no_grad():
            model.generate(**inp, max_new_tokens=3, generation_config=gen_cfg)
    finally:
        patched._lm = orig_lm
    full_pos, delta = patched.mrope_position_ids(inp["input_ids"], VIDEO_ID, inp["video_grid_thw"], 2)
    decode = [p for p in seen if p.shape[-1] == 1]
    assert len(decode) == 2                                         # 3 new tokens = prefill + 2 decode steps
    want_first = int(full_pos.max()) + 1                            # = S_orig + rope_delta
    assert want_first == inp["input_ids"].shape[1] + int(delta)
    prefill_last = max(int(p.max()) for p in seen if p.shape[-1] > 1)
    for i, p in enumerate(decode):
        assert p.shape == (3, 1, 1) and p.flatten().tolist() == [want_first + i] * 3
    assert int(decode[0].min()) > prefill_last
