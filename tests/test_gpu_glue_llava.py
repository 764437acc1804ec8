"""GPU: the transformers-5 LLaVA-OneVision glue (retake/llava_onevision.py) on a tiny random-init model."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
VIDEO_ID, IMAGE_ID = 900, 901


def tiny_model():
    from transformers import LlavaOnevisionConfig, LlavaOnevisionForConditionalGeneration
    cfg = LlavaOnevisionConfig(
        vision_config=dict(model_type="siglip_vision_model", hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                           num_attention_heads=4, image_size=56, patch_size=14),
        text_config=dict(model_type="qwen2", hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                         num_attention_heads=4, num_key_value_heads=2, vocab_size=1000, max_position_embeddings=4096),
        video_token_index=VIDEO_ID, image_token_index=IMAGE_ID, vision_feature_layer=-1,
        vision_feature_select_strategy="full")
    cfg._attn_implementation = "sdpa"
    torch.manual_seed(0)
    return LlavaOnevisionForConditionalGeneration(cfg).to(torch.bfloat16).cuda().eval()


def make_inputs(T=16, pre=5, post=6):
    g = torch.Generator().manual_seed(1)
    n_vid = T * 4 + 1
    ids = torch.cat([torch.randint(0, 800, (pre,), generator=g), torch.full((n_vid,), VIDEO_ID),
                     torch.randint(0, 800, (post,), generator=g)])
    scene = torch.randn(1, 3, 56, 56, generator=g)
    px = scene.repeat(T, 1, 1, 1) + 0.3 * torch.randn(T, 3, 56, 56, generator=g)
    px[T // 2:] += torch.randn(1, 3, 56, 56, generator=g)
    return dict(input_ids=ids[None].cuda(), pixel_values_videos=px[None].to(torch.bfloat16).cuda(),
                attention_mask=torch.ones(1, ids.numel(), dtype=torch.long).cuda())


def lv_kwargs(vis=True, rv=1.0, kv=True, rkv=1.0, reforge=False, chunk_frames=4, deferred=False):
    return {"frame_chunk_size": 8, "chunked_prefill_frames": chunk_frames, "visual_compression": vis,
            "visual_compression_kwargs": {"compression_ratio": rv, "compression_method": "Keyframe", "patch_sync": False,
                                          "return_keyframe_mask": True},
            "kvcache_compression": kv,
            "kvcache_compression_kwargs": {"dynamic_compression_ratio": False, "compression_ratio": rkv,
                                           "compression_method": "pivotkv", "pos_embed_reforge": reforge,
                                           "deferred_compression": deferred}}


@pytest.fixture()
def patched():
    from retake import llava_onevision, monkeypatch
    monkeypatch.patch_llava_onevision("retake")
    yield llava_onevision
    llava_onevision.uninstall()


def test_chunked_prefill_without_compression_equals_stock(patched):
    from retake import monkeypatch
    model = tiny_model()
    monkeypatch.patch_llava_onevision_config(model.config, {"longvideo_kwargs": lv_kwargs(vis=False, kv=False)})
    inp = make_inputs()
    assert model.model.get_chunk_size(model.config, inp["pixel_values_videos"]) == 4 * 2 * 2
    assert [k for _, _, k in model.model.segment_input_ids(inp["input_ids"])] == ["text", "video", "text"]
    with torch.no_grad():
        a = model(**inp, use_cache=True).logits[0, -1].float()
        model.config.longvideo_kwargs = lv_kwargs(vis=False, kv=False, chunk_frames=1000)
        b = model(**inp, use_cache=True).logits[0, -1].float()
        patched.uninstall()
        model.config.longvideo_kwargs = None
        c = model(**inp, use_cache=True).logits[0, -1].float()          # stock transformers forward
        patched.install()
    assert torch.allclose(a, b, atol=0.08, rtol=0.05) and torch.allclose(a, c, atol=0.08, rtol=0.05)


def test_deferred_compression_gives_the_same_prefill(patched):
    """1-D positions (Qwen2 rotary): one batched compression per chunk at after_forward() == compression inside update()"""
    model = tiny_model()
    inp = make_inputs()
    outs = []
    for deferred in (False, True):
        model.config.longvideo_kwargs = lv_kwargs(vis=True, rv=0.5, kv=True, rkv=0.5, reforge=True, deferred=deferred)
        with torch.no_grad():
            outs.append(model(**inp, use_cache=True))
    a, b = outs
    assert b.past_key_values.deferred_compression and not a.past_key_values.deferred_compression
    assert torch.equal(a.logits, b.logits)
    for l in range(2):
        assert torch.equal(a.past_key_values.layers[l].keys, b.past_key_values.layers[l].keys)
        assert torch.equal(a.past_key_values.layers[l].values, b.past_key_values.layers[l].values)
        assert torch.equal(a.past_key_values.position_cache[l], b.past_key_values.position_cache[l])


@pytest.mark.parametrize("reforge", [False, True])
def test_compressed_prefill_and_generate(patched, reforge):
    from retake.longvideo_cache import PivotKVCache
    model = tiny_model()
    model.config.longvideo_kwargs = lv_kwargs(vis=True, rv=0.5, kv=True, rkv=0.5, reforge=reforge)
    inp = make_inputs()
    with torch.no_grad():
        o = model(**inp, use_cache=True)
    cache = o.past_key_values
    assert isinstance(cache, PivotKVCache)
    # DPSelect keeps 8 of 16 frames -> 32 video slots (the newline slot is dropped, reference quirk) in two 16-token
    # chunks; PivotKV keeps 8 of each
    want_len = 5 + 2 * 8 + 6
    assert [cache.get_seq_length(l) for l in range(2)] == [want_len, want_len] and cache.num_evicted_tokens == [16, 16]
    if reforge:
        t = cache.position_cache[0][0]
        assert t.shape[0] == want_len and bool((t[1:] >= t[:-1]).all())
    assert torch.isfinite(o.logits.float()).all()
    gen_cfg = copy.deepcopy(model.generation_config)
    gen_cfg.do_sample = False
    with torch.no_grad():
        out_ids = model.generate(**inp, max_new_tokens=4, generation_config=gen_cfg)
    assert out_ids.shape[1] == inp["input_ids"].shape[1] + 4


@pytest.mark.parametrize("method", ["MA-LLM", "MA-LLM-hard"])
def test_mallm_visual_compression_methods(patched, method):
    """`compression_method: MA-LLM / MA-LLM-hard` (llava_onevision.py:235-243) through the fused loop"""
    model = tiny_model()
    kw = lv_kwargs(vis=True, rv=0.5, kv=True, rkv=0.5, reforge=False)
    kw["visual_compression_kwargs"]["compression_method"] = method
    model.config.longvideo_kwargs = kw
    with torch.no_grad():
        o = model(**make_inputs(), use_cache=True)
    want_len = 5 + 2 * 8 + 6
    assert [o.past_key_values.get_seq_length(l) for l in range(2)] == [want_len, want_len]
    assert torch.isfinite(o.logits.float()).all()
