"""Model-level glue for LLaVA-OneVision / LLaVA-Video (Qwen2 LLM + SigLIP), re-targeted at transformers 5.x
(row f1 of SURVEY.md section 8; reference ``retake/llava_onevision.py``, ``retake/monkeypatch.py:67-77``).

* ``retake_Qwen2Attention_forward`` (reference ``:59-141``): when ``pos_embed_reforge`` is on the 1-D position ids of the
  chunk are re-based per layer on that layer's (compacted) cache and the rotary tables recomputed; the cache receives
  ``query_states / position_ids / rotary_emb`` through ``cache_kwargs``.  transformers 5.x keeps the rotary module on the
  text model, so it travels on the cache object (the reference adds one per attention layer in a patched ``__init__``,
  ``:48-56``; ``retake_Qwen2Attention_init`` is kept as a no-op wrapper for API compatibility).
* ``get_chunk_size`` (``:144-160``), ``segment_input_ids`` (``:163-198``), ``compress_video_tokens`` (``:201-269``, DPSelect on
  the SigLIP hidden states BEFORE the projector), ``forge_input_chunks`` (``:272-303``) keep their names.
* ``retake_LlavaOnevisionModel_forward`` (``:306-583``): frame-chunked vision tower, DPSelect, projector + 2x bilinear
  pooling + ``image_newline``, chunked prefill with the same cache hooks as the reference loop (``:487-538``).

Reference quirks kept on purpose: the video span of ``input_ids`` is cut to ``t * 196`` slots so the trailing
``image_newline`` feature is dropped whenever visual compression is on (``:245-252``); the key-patch mask has ``t * 729``
entries of which only the first ``t * 196`` land on token slots (``:485-486``).
"""
from __future__ import annotations

import math

import torch
from transformers.models.llava_onevision import modeling_llava_onevision as hf
from transformers.models.qwen2 import modeling_qwen2 as hfq

from .longvideo_cache import PivotKVCache, build_kvcache
from .visual_compression import mallm_compress, memory_bank_compress_keyframe

__all__ = ["install", "uninstall", "retake_Qwen2Attention_init", "retake_Qwen2Attention_forward",
           "retake_LlavaOnevisionModel_forward",
           "retake_LlavaOnevisionForConditionalGeneration_get_chunk_size",
           "retake_LlavaOnevisionForConditionalGeneration_segment_input_ids",
           "retake_LlavaOnevisionForConditionalGeneration_compress_video_tokens",
           "retake_LlavaOnevisionForConditionalGeneration_forge_input_chunks"]

_ORIG = {}
POOL_STRIDE = 2


def retake_Qwen2Attention_init(self, config, layer_idx=None):
    _ORIG.get("attention_init", hfq.Qwen2Attention.__init__)(self, config, layer_idx)


def retake_Qwen2Attention_forward(self, hidden_states, position_embeddings, attention_mask, past_key_values=None,
                                  **kwargs):
    input_shape = hidden_states.shape[:-1]
    hidden_shape = (*input_shape, -1, self.head_dim)
    query_states = self.q_proj(hidden_states).view(hidden_shape).transpose(1, 2)
    key_states = self.k_proj(hidden_states).view(hidden_shape).transpose(1, 2)
    value_states = self.v_proj(hidden_states).view(hidden_shape).transpose(1, 2)

    cache = past_key_values
    pos = getattr(cache, "retake_position_ids", None)
    rotary = getattr(cache, "retake_rotary_emb", None)
    position_ids = None
    if isinstance(cache, PivotKVCache) and cache.pos_embed_reforge and pos is not None and rotary is not None:
        # re-base this chunk's positions on the layer's compacted cache (sync-free form of reference :80-89)
        pos_all = getattr(cache, "retake_position_ids_all", None)
        if pos_all is not None and self.layer_idx < pos_all.shape[0]:
            position_ids = pos_all[self.layer_idx]          # all layers re-based up front by _lm()
        else:
            prev = cache.get_prev_temporal_idx(self.layer_idx)
            position_ids = pos.clone()
            position_ids[0, :] += prev + 1 - pos[0, 0]
        cos, sin = rotary(value_states, position_ids)
    else:
        cos, sin = position_embeddings
        position_ids = pos
    query_states, key_states = hfq.apply_rotary_pos_emb(query_states, key_states, cos, sin)

    if cache is not None:
        if isinstance(cache, PivotKVCache):
            cache_kwargs = {"sin": sin, "cos": cos, "query_states": query_states, "position_ids": position_ids,
                            "rotary_emb": rotary,
                            # with re-forging the ids above are this layer's own copy: the cache may keep them as they are
                            "position_ids_owned": position_ids is not None and position_ids is not pos}
            key_states, value_states = cache.update(key_states, value_states, self.layer_idx, cache_kwargs)
        else:
            key_states, value_states = cache.update(key_states, value_states, self.layer_idx)

    attention_interface = hfq.ALL_ATTENTION_FUNCTIONS.get_interface(self.config._attn_implementation,
                                                                    hfq.eager_attention_forward)
    attn_output, attn_weights = attention_interface(
        self, query_states, key_states, value_states, attention_mask,
        dropout=0.0 if not self.training else self.attention_dropout, scaling=self.scaling,
        sliding_window=self.sliding_window, **kwargs)
    attn_output = attn_output.reshape(*input_shape, -1).contiguous()
    return self.o_proj(attn_output), attn_weights


# ------------------------------------------------------------------------------------------------- helpers
def retake_LlavaOnevisionForConditionalGeneration_get_chunk_size(self, config, pixel_values_videos):
    lv = getattr(config, "longvideo_kwargs", None)
    chunk_frames = lv.get("chunked_prefill_frames", None) if lv else None
    if chunk_frames is None or pixel_values_videos is None:
        return None
    T, _, H, W = pixel_values_videos[0].shape
    patch = self.config.vision_config.patch_size
    H = math.ceil(H // patch / POOL_STRIDE)
    W = math.ceil(W // patch / POOL_STRIDE)
    return min(chunk_frames, T) * H * W


def retake_LlavaOnevisionForConditionalGeneration_segment_input_ids(self, input_ids):
    video_id = getattr(self.config, "video_token_id", None)
    if video_id is None:
        video_id = self.config.video_token_index
    is_video = (input_ids[0] == video_id).tolist()
    segments, start = [], 0
    for i in range(1, len(is_video) + 1):
        if i == len(is_video) or is_video[i] != is_video[start]:
            segments.append((start, i, "video" if is_video[start] else "text"))
            start = i
    return segments


def retake_LlavaOnevisionForConditionalGeneration_compress_video_tokens(self, input_ids=None, attention_mask=None,
                                                                       selected_video_feature=None, position_ids=None,
                                                                       cache_position=None, labels=None):
    """DPSelect on the SigLIP features ``[T, N, C]`` and the matching truncation (reference ``:201-269``)."""
    lv = getattr(self.config, "longvideo_kwargs", None) or {}
    grid_t, grid_hw = selected_video_feature.shape[:2]
    tgt_grid_t, keypatches_mask = grid_t, None
    if lv.get("visual_compression", False):
        kw = lv["visual_compression_kwargs"]
        assert labels is None
        assert input_ids.shape[0] == 1, "Currently, only inference are supported"
        video_id = getattr(self.config, "video_token_id", None)
        if video_id is None:
            video_id = self.config.video_token_index
        idx = torch.where(input_ids[0] == video_id)[0]
        s_index, e_index = int(idx[0]), int(idx[-1])
        side = self.config.vision_config.image_size // self.config.vision_config.patch_size
        grid_hw_after_pool = math.ceil(side / POOL_STRIDE) ** 2
        ori_seq_len = input_ids.shape[1]
        tgt_grid_t = max(1, round(kw.get("compression_ratio") * grid_t))
        bank = selected_video_feature.reshape(1, grid_t, grid_hw, -1)
        if kw.get("compression_method") == "Keyframe":
            bank, keypatches_mask = memory_bank_compress_keyframe(bank, tgt_grid_t, 3, sync=kw.get("patch_sync"))
            keypatches_mask = keypatches_mask if kw.get("return_keyframe_mask") else None
        elif kw.get("compression_method") in ("MA-LLM", "MA-LLM-hard"):   # llava_onevision.py:235-243, fused
            bank, _ = mallm_compress(bank, tgt_grid_t, sync=bool(kw.get("patch_sync")),
                                     hard=kw.get("compression_method") == "MA-LLM-hard")
        else:
            raise NotImplementedError(f"unknown visual compression method {kw.get('compression_method')!r}")
        selected_video_feature = bank[0]
        mem_len_after = tgt_grid_t * grid_hw_after_pool
        input_ids = torch.cat([input_ids[:, :s_index], input_ids[:, s_index:e_index + 1][:, :mem_len_after],
                               input_ids[:, e_index + 1:]], dim=1)
        num_token_diff = ori_seq_len - input_ids.shape[1]
        if num_token_diff and attention_mask is not None:
            attention_mask = attention_mask[:, num_token_diff:]
        if num_token_diff and position_ids is not None:
            position_ids = position_ids[:, :-num_token_diff]
        if num_token_diff and cache_position is not None:
            cache_position = cache_position[:-num_token_diff]
    return input_ids, attention_mask, selected_video_feature, position_ids, cache_position, tgt_grid_t, keypatches_mask


def retake_LlavaOnevisionForConditionalGeneration_forge_input_chunks(self, ss, ee, modality_segments, position_ids,
                                                                    cache_position, attention_mask, past_key_values,
                                                                    inputs_embeds):
    lv = getattr(self.config, "longvideo_kwargs", None) or {}
    kw = lv.get("kvcache_compression_kwargs", {}) if lv.get("kvcache_compression", False) else {}
    if kw.get("prompt_guided_compression", False) and kw.get("compression_ratio", 1) < 1.0:
        raise NotImplementedError("prompt_guided_compression is not supported")
    cache_position_chunk = cache_position[:ee] if cache_position is not None else None
    attention_mask_chunk = attention_mask[:, :ee] if attention_mask is not None else None
    return position_ids[:, ss:ee], cache_position_chunk, attention_mask_chunk, inputs_embeds[:, ss:ee], None


def _siglip_features(self, pixel_values_videos, vision_feature_layer):
    """[T, N, C] hidden states of the selected vision layer, vision tower run over frame chunks (reference :420-436)"""
    lv = getattr(self.config, "longvideo_kwargs", None) or {}
    frame_chunk = lv.get("frame_chunk_size", 1_000_000_000)
    b, frames, ch, h, w = pixel_values_videos.shape
    px = pixel_values_videos.view(b * frames, ch, h, w)

    def run(p):
        out = self.vision_tower(p, output_hidden_states=True, return_dict=True)
        if isinstance(vision_feature_layer, int):
            return out.hidden_states[vision_feature_layer]
        return torch.cat([out.hidden_states[i] for i in vision_feature_layer], dim=-1)

    if b * frames < frame_chunk:
        return run(px)
    return torch.cat([run(px[i:i + frame_chunk]) for i in range(0, b * frames, frame_chunk)])


def _lm(self, cache, inputs_embeds, position_ids, **kwargs):
    cache.retake_position_ids = position_ids if isinstance(cache, PivotKVCache) else None
    cache.retake_position_ids_all = None
    if isinstance(cache, PivotKVCache) and cache.pos_embed_reforge and position_ids is not None:
        cache.retake_position_ids_all = cache.rebased_position_ids(position_ids, self.language_model.config.num_hidden_layers)
    cache.retake_rotary_emb = self.language_model.rotary_emb
    return self.language_model(attention_mask=None, position_ids=position_ids, past_key_values=cache,
                               inputs_embeds=inputs_embeds, use_cache=True, **kwargs)


def retake_LlavaOnevisionModel_forward(self, input_ids=None, pixel_values=None, image_sizes=None, pixel_values_videos=None,
                                       image_sizes_videos=None, attention_mask=None, position_ids=None,
                                       past_key_values=None, inputs_embeds=None, vision_feature_layer=None,
                                       vision_feature_select_strategy=None, vision_aspect_ratio=None, batch_num_images=None,
                                       use_cache=None, **kwargs):
    lv = getattr(self.config, "longvideo_kwargs", None)
    chunk_size = None
    if lv and input_ids is not None and input_ids.shape[1] > 1 and pixel_values_videos is not None and pixel_values is None:
        chunk_size = self.get_chunk_size(self.config, pixel_values_videos)

    if chunk_size is None:
        if isinstance(past_key_values, PivotKVCache) and input_ids is not None:
            assert input_ids.shape[0] == 1
            q_len = input_ids.shape[1]
            seen = past_key_values.retake_seen_tokens
            pos = (torch.arange(q_len, device=input_ids.device) + seen)[None]
            past_key_values.retake_seen_tokens = seen + q_len
            past_key_values.kvcache_compression = False
            out = _lm(self, past_key_values, self.get_input_embeddings()(input_ids), pos, **kwargs)
            return hf.LlavaOnevisionModelOutputWithPast(last_hidden_state=out.last_hidden_state,
                                                        past_key_values=past_key_values)
        return _ORIG["model_forward"](self, input_ids=input_ids, pixel_values=pixel_values, image_sizes=image_sizes,
                                      pixel_values_videos=pixel_values_videos, image_sizes_videos=image_sizes_videos,
                                      attention_mask=attention_mask, position_ids=position_ids,
                                      past_key_values=past_key_values, inputs_embeds=inputs_embeds,
                                      vision_feature_layer=vision_feature_layer,
                                      vision_feature_select_strategy=vision_feature_select_strategy,
                                      vision_aspect_ratio=vision_aspect_ratio, batch_num_images=batch_num_images,
                                      use_cache=use_cache, **kwargs)

    assert input_ids.shape[0] == 1, "Batch inference of long video is not supported yet!"
    if attention_mask is not None and not bool(attention_mask.all()):
        raise NotImplementedError("padded inputs are not supported (batch size 1, no padding)")
    if lv.get("kvcache_compression", False):
        kw = lv["kvcache_compression_kwargs"]
        if kw.get("dynamic_compression_ratio", False):
            input_length, max_len = input_ids.shape[1], kw["max_input_length"]
            kw["compression_ratio"] = 1 if input_length <= max_len else max_len / input_length
    cache = build_kvcache(self.config)
    vision_feature_layer = vision_feature_layer if vision_feature_layer is not None else self.config.vision_feature_layer
    strategy = (vision_feature_select_strategy if vision_feature_select_strategy is not None
                else self.config.vision_feature_select_strategy)

    batch_size, frames = pixel_values_videos.shape[:2]
    prompt_tokens = input_ids.shape[1]        # BEFORE visual compression: what generate() derives decode positions from
    position_ids = torch.arange(input_ids.shape[1], device=input_ids.device)[None]
    feats = _siglip_features(self, pixel_values_videos, vision_feature_layer)                 # [T, N, C]
    input_ids, attention_mask, feats, position_ids, _, frames, keypatches_mask = self.compress_video_tokens(
        input_ids=input_ids, attention_mask=attention_mask, selected_video_feature=feats, position_ids=position_ids,
        cache_position=None, labels=None)
    if strategy == "default":
        feats = feats[:, 1:]
    video_features = self.apply_pooling(self.multi_modal_projector(feats))
    video_features = video_features.reshape(batch_size, frames * video_features.shape[1], -1)
    newline = self.image_newline[None, None, :].repeat(batch_size, 1, 1).to(video_features.device)
    video_features = torch.cat((video_features, newline), dim=1).flatten(0, 1)

    inputs_embeds = self.get_input_embeddings()(input_ids)
    video_id = getattr(self.config, "video_token_id", None)
    if video_id is None:
        video_id = self.config.video_token_index
    video_mask = input_ids == video_id
    n_slots = int(video_mask.sum())
    if n_slots > video_features.shape[0]:
        raise ValueError(f"Video features and video tokens do not match: tokens: {n_slots}, features {video_features.shape[0]}")
    inputs_embeds = inputs_embeds.masked_scatter(video_mask.unsqueeze(-1).expand_as(inputs_embeds),
                                                 video_features.to(inputs_embeds.device, inputs_embeds.dtype))
    if keypatches_mask is not None:
        keypatches_mask = torch.zeros_like(input_ids).bool().masked_scatter(video_mask, keypatches_mask)

    modality_segments = self.segment_input_ids(input_ids)
    compress = getattr(cache, "kvcache_compression", False)
    outputs = None
    for s, e, kind in modality_segments:
        if kind == "text":
            cache.kvcache_compression = False
            outputs = _lm(self, cache, inputs_embeds[:, s:e], position_ids[:, s:e], **kwargs)
        else:
            cache.kvcache_compression = compress
            for c in range(math.ceil((e - s) / chunk_size)):
                ss, ee = s + c * chunk_size, min(s + (c + 1) * chunk_size, e)
                if keypatches_mask is not None:
                    cache.keypatches_mask_chunk = keypatches_mask[0, ss:ee]
                pos_chunk, _, _, emb_chunk, prompt_length = self.forge_input_chunks(
                    ss, ee, modality_segments, position_ids, None, None, cache, inputs_embeds)
                if hasattr(cache, "before_forward"):
                    cache.before_forward(prompt_length=prompt_length)
                outputs = _lm(self, cache, emb_chunk, pos_chunk, **kwargs)
                if hasattr(cache, "after_forward"):
                    cache.after_forward()
            cache.keypatches_mask_chunk = None
            cache.kvcache_compression = False
    # generate() numbers decoded tokens from the ORIGINAL prompt length (attention-mask cumsum / cache_position), not from
    # the sequence DPSelect / MA-LLM shortened (reference llava_onevision.py:262-265 only trims the prefill ids)
    cache.retake_seen_tokens = prompt_tokens
    return hf.LlavaOnevisionModelOutputWithPast(last_hidden_state=outputs.last_hidden_state, past_key_values=cache)


def install():
    """attach the ReTaKe forwards to the transformers 5.x classes (``monkeypatch.py:67-77``)"""
    if "model_forward" not in _ORIG:
        _ORIG["model_forward"] = hf.LlavaOnevisionModel.forward
        _ORIG["attention_forward"] = hfq.Qwen2Attention.forward
        _ORIG["attention_init"] = hfq.Qwen2Attention.__init__
    hfq.Qwen2Attention.forward = retake_Qwen2Attention_forward
    for cls in (hf.LlavaOnevisionModel, hf.LlavaOnevisionForConditionalGeneration):
        cls.get_chunk_size = retake_LlavaOnevisionForConditionalGeneration_get_chunk_size
        cls.segment_input_ids = retake_LlavaOnevisionForConditionalGeneration_segment_input_ids
        cls.compress_video_tokens = retake_LlavaOnevisionForConditionalGeneration_compress_video_tokens
        cls.forge_input_chunks = retake_LlavaOnevisionForConditionalGeneration_forge_input_chunks
    hf.LlavaOnevisionModel.forward = retake_LlavaOnevisionModel_forward


def uninstall():
    if _ORIG:
        hf.LlavaOnevisionModel.forward = _ORIG["model_forward"]
        hfq.Qwen2Attention.forward = _ORIG["attention_forward"]
