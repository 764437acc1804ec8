"""Model-level glue for Qwen2-VL, re-targeted at transformers 5.x (row f1 of SURVEY.md section 8).

The reference patches transformers 4.48 classes (``retake/qwen2_vl.py``, ``retake/monkeypatch.py:51-62``); three of
them no longer exist in 5.x (one ``Qwen2VLAttention`` with pluggable attention functions replaces the eager / SDPA /
FA2 subclasses, the rotary module moved from the attention layer to the text model, caches store ``layers[i].keys``).
This module keeps the reference's flow and hook order and re-attaches it to the 5.x classes:

* ``retake_Qwen2VLAttention_forward``  - ``qwen2_vl.py:40-363``: per-layer temporal re-basing of the position ids when
  ``pos_embed_reforge`` is on (``:68-73``), rotary, ``past_key_values.update(k, v, layer, cache_kwargs)`` with
  ``query_states / position_ids / rotary_emb / mrope_section`` (``:297-301``), then the stock attention function.
* ``compress_video_tokens`` (``:366-442``), ``segment_input_ids`` (``:444-475``), ``get_chunk_size`` (``:477-491``),
  ``forge_input_chunks`` (``:493-519``) - same names, bound to ``Qwen2VLModel`` (and reachable from
  ``Qwen2VLForConditionalGeneration``).
* ``retake_Qwen2VLModel_forward`` - ``qwen2_vl.py:522-733``: frame-chunked vision tower, DPSelect, chunked prefill over
  text / video segments with ``kvcache_compression`` / ``keypatches_mask_chunk`` / ``before_forward`` /
  ``after_forward`` driven exactly like the reference loop (``:670-718``), decode with positions counted in
  uncompressed tokens.

Differences that are deliberate: position ids are computed here with the Qwen2-VL M-RoPE rule the reference was written
against (temporal index advances by one per temporal grid; transformers 5.5's ``get_rope_index`` keeps it constant
within a video); modality segments are taken AFTER DPSelect shortened the sequence (the reference takes them before,
``:559`` vs ``:619``, which only works for the shipped ratio 1.0); an all-ones attention mask is dropped so that the
causal mask follows the compressed cache length.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from transformers.models.qwen2_vl import modeling_qwen2_vl as hf

from .longvideo_cache import PivotKVCache, build_kvcache
from .visual_compression import mallm_compress, memory_bank_compress_keyframe

__all__ = ["install", "retake_Qwen2VLAttention_forward", "retake_Qwen2VLModel_forward",
           "retake_Qwen2VLForConditionalGeneration_compress_video_tokens",
           "retake_Qwen2VLForConditionalGeneration_segment_input_ids",
           "retake_Qwen2VLForConditionalGeneration_get_chunk_size",
           "retake_Qwen2VLForConditionalGeneration_forge_input_chunks", "mrope_position_ids"]

_ORIG = {}


# ------------------------------------------------------------------------------------------------ attention
def retake_Qwen2VLAttention_forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_values=None,
                                    output_attentions=False, use_cache=False, position_embeddings=None, **kwargs):
    bsz, q_len, _ = hidden_states.size()
    query_states = self.q_proj(hidden_states).view(bsz, q_len, -1, self.head_dim).transpose(1, 2)
    key_states = self.k_proj(hidden_states).view(bsz, q_len, -1, self.head_dim).transpose(1, 2)
    value_states = self.v_proj(hidden_states).view(bsz, q_len, -1, self.head_dim).transpose(1, 2)
    mrope_section = self.config.rope_parameters["mrope_section"]

    cache = past_key_values
    pos3d = getattr(cache, "retake_position_ids", None)
    rotary = getattr(cache, "retake_rotary_emb", None)
    if isinstance(cache, PivotKVCache) and cache.pos_embed_reforge and pos3d is not None and rotary is not None:
        # temporal ids continue after this layer's (compacted) cache; sync-free form of qwen2_vl.py:68-73
        assert bsz == 1
        # Every layer gets ITS OWN id tensor (handed over to the cache: deferred compression reads it at after_forward()).
        # The re-basing only depends on the layer's cache as the previous call left it, so _lm() has computed the ids of
        # all layers up front (PivotKVCache.rebased_position_ids: a handful of launches per call instead of a few per layer).
        pos_all = getattr(cache, "retake_position_ids_all", None)
        if pos_all is not None and self.layer_idx < pos_all.shape[0]:
            pos3d = pos_all[self.layer_idx]
        else:
            prev = cache.get_prev_temporal_idx(self.layer_idx)
            shifted = pos3d.clone()
            shifted[0, 0, :] += prev + 1 - pos3d[0, 0, 0]
            pos3d = cache.retake_position_ids = shifted
        cos, sin = rotary(value_states, pos3d)
    else:
        cos, sin = position_embeddings
    query_states, key_states = hf.apply_multimodal_rotary_pos_emb(query_states, key_states, cos, sin, mrope_section)

    if cache is not None:
        if isinstance(cache, PivotKVCache):
            cache_kwargs = {"sin": sin, "cos": cos, "query_states": query_states, "position_ids": pos3d,
                            "rotary_emb": rotary, "mrope_section": mrope_section,
                            "position_ids_owned": cache.pos_embed_reforge}
            key_states, value_states = cache.update(key_states, value_states, self.layer_idx, cache_kwargs)
        else:
            key_states, value_states = cache.update(key_states, value_states, self.layer_idx)

    attention_interface = hf.ALL_ATTENTION_FUNCTIONS.get_interface(self.config._attn_implementation,
                                                                   hf.eager_attention_forward)
    attn_output, attn_weights = attention_interface(
        self, query_states, key_states, value_states, attention_mask,
        dropout=0.0 if not self.training else self.attention_dropout, scaling=self.scaling,
        sliding_window=self.sliding_window, position_ids=position_ids, **kwargs)
    attn_output = attn_output.reshape(bsz, q_len, -1).contiguous()
    return self.o_proj(attn_output), attn_weights


# ------------------------------------------------------------------------------------------- model helpers
def mrope_position_ids(input_ids: torch.Tensor, video_token_id: int, video_grid_thw: Optional[torch.Tensor],
                       spatial_merge_size: int):
    """Qwen2-VL M-RoPE ids ``[3, 1, S]`` for ONE sequence with at most one video (the reference's supported case,
    ``qwen2_vl.py:389-390``): text advances all three rows together, a video block of ``T x H x W`` merged tokens gets
    (t, h, w) offsets from the block start, and the text after it continues at max + 1.  Returns (ids, rope_delta)."""
    ids = input_ids[0]
    S = ids.numel()
    dev = ids.device
    vid = torch.nonzero(ids == video_token_id)[:, 0]
    if vid.numel() == 0 or video_grid_thw is None:
        pos = torch.arange(S, device=dev).view(1, 1, -1).expand(3, 1, -1).clone()
        return pos, torch.zeros(1, 1, dtype=torch.long, device=dev)
    s, e = int(vid[0]), int(vid[-1]) + 1
    T, H, W = (int(v) for v in video_grid_thw[0])
    H, W = H // spatial_merge_size, W // spatial_merge_size
    n = T * H * W
    assert e - s == n, "video tokens do not match video_grid_thw"
    ar = torch.arange(n, device=dev)
    vis = torch.stack([ar // (H * W), (ar // W) % H, ar % W]) + s
    after = s + max(T, H, W)
    pos = torch.cat([torch.arange(s, device=dev).expand(3, -1), vis,
                     (torch.arange(S - e, device=dev) + after).expand(3, -1)], dim=1)[:, None]
    delta = (pos.max() + 1 - S).view(1, 1)
    return pos.contiguous(), delta


def retake_Qwen2VLForConditionalGeneration_compress_video_tokens(self, input_ids=None, attention_mask=None,
                                                                 video_embeds=None, cache_position=None,
                                                                 position_ids=None, labels=None, video_grid_thw=None):
    """DPSelect on the video embeddings and the matching truncation of ids / mask / positions (``qwen2_vl.py:366-442``)."""
    lv = getattr(self.config, "longvideo_kwargs", None) or {}
    keypatches_mask = None
    if lv.get("visual_compression", False):
        kw = lv["visual_compression_kwargs"]
        ratio, method = kw.get("compression_ratio"), kw.get("compression_method")
        assert labels is None
        assert video_grid_thw.shape[0] <= 1, "Currently, interleaved videos are not supported"
        assert input_ids.shape[0] == 1, "Currently, only inference are supported"
        video_token_id = self.config.video_token_id
        idx = torch.where(input_ids[0] == video_token_id)[0]
        s_index, e_index = int(idx[0]), int(idx[-1])
        grid_t = int(video_grid_thw[0][0])
        grid_hw = video_embeds.shape[0] // grid_t
        ori_seq_len = input_ids.shape[1]
        tgt_mem_len = max(1, round(ratio * grid_t))
        num_frame_diff = grid_t - tgt_mem_len
        bank = video_embeds.reshape(1, grid_t, grid_hw, -1)
        if method == "Keyframe":
            bank, keypatches_mask = memory_bank_compress_keyframe(bank, tgt_mem_len, 3, sync=kw.get("patch_sync"))
            keypatches_mask = keypatches_mask if kw.get("return_keyframe_mask") else None
        elif method in ("MA-LLM", "MA-LLM-hard"):                   # the reference's while-loop (qwen2_vl.py:402-409), fused
            bank, _ = mallm_compress(bank, tgt_mem_len, sync=bool(kw.get("patch_sync")), hard=method == "MA-LLM-hard")
        else:
            raise NotImplementedError(f"unknown visual compression method {method!r}")
        video_embeds = bank.flatten(1, 2)[0]
        tgt_seq_len = video_embeds.shape[0]
        input_ids = torch.cat([input_ids[:, :s_index], input_ids[:, s_index:e_index + 1][:, :tgt_seq_len],
                               input_ids[:, e_index + 1:]], dim=1)
        num_token_diff = ori_seq_len - input_ids.shape[1]
        if num_token_diff and attention_mask is not None:
            attention_mask = attention_mask[:, :-num_token_diff]
        if num_token_diff and cache_position is not None:
            cache_position = cache_position[:-num_token_diff]
        if position_ids is not None:
            position_ids = torch.cat([position_ids[..., :s_index], position_ids[..., s_index:e_index + 1][..., :tgt_seq_len],
                                      position_ids[..., e_index + 1:]], dim=2)
            position_ids[:, :, s_index + tgt_seq_len:] -= num_frame_diff
    return input_ids, attention_mask, video_embeds, cache_position, position_ids, labels, keypatches_mask


def retake_Qwen2VLForConditionalGeneration_segment_input_ids(self, input_ids):
    """[(s, e, 'video' | 'text')] covering the sequence in order (``qwen2_vl.py:444-475``)."""
    is_video = (input_ids[0] == self.config.video_token_id).tolist()
    segments, start = [], 0
    for i in range(1, len(is_video) + 1):
        if i == len(is_video) or is_video[i] != is_video[start]:
            segments.append((start, i, "video" if is_video[start] else "text"))
            start = i
    return segments


def retake_Qwen2VLForConditionalGeneration_get_chunk_size(self, config, video_grid_thw):
    """tokens per prefill chunk = min(chunked_prefill_frames, T) * H * W / (merge^2 * temporal_patch) (``:477-491``)."""
    lv = getattr(config, "longvideo_kwargs", None)
    chunk_frames = lv.get("chunked_prefill_frames", None) if lv else None
    if chunk_frames is None or video_grid_thw is None:
        return None
    T, H, W = (int(v) for v in video_grid_thw[0])
    t_factor = config.vision_config.spatial_merge_size ** 2 * config.vision_config.temporal_patch_size
    return min(chunk_frames, T) * H * W // t_factor


def retake_Qwen2VLForConditionalGeneration_forge_input_chunks(self, ss, ee, modality_segments, cache_position,
                                                              position_ids, attention_mask, past_key_values,
                                                              inputs_embeds):
    """slices of one video chunk (``qwen2_vl.py:493-519``); prompt-guided compression is plumbed but dead in the
    reference (``longvideo_cache.py:146-147``) and not offered here"""
    lv = getattr(self.config, "longvideo_kwargs", None) or {}
    kw = lv.get("kvcache_compression_kwargs", {}) if lv.get("kvcache_compression", False) else {}
    if kw.get("prompt_guided_compression", False) and kw.get("compression_ratio", 1) < 1.0:
        raise NotImplementedError("prompt_guided_compression is not supported")
    cache_position_chunk = cache_position[:ee] if cache_position is not None else None
    attention_mask_chunk = attention_mask[:, :ee] if attention_mask is not None else None
    return cache_position_chunk, position_ids[:, :, ss:ee], attention_mask_chunk, inputs_embeds[:, ss:ee], None


def _video_features(self, pixel_values_videos, video_grid_thw):
    """vision tower over frame chunks of ``frame_chunk_size`` temporal grids (``qwen2_vl.py:597-617``)"""
    lv = getattr(self.config, "longvideo_kwargs", None) or {}
    frame_chunk = lv.get("frame_chunk_size", 1_000_000_000)
    pixel_values_videos = pixel_values_videos.type(self.visual.dtype)
    grid_t, grid_h, grid_w = (int(v) for v in video_grid_thw[0])
    if grid_t < frame_chunk:
        return self.visual(pixel_values_videos, grid_thw=video_grid_thw).pooler_output
    d = pixel_values_videos.shape[-1]
    pv = pixel_values_videos.reshape(grid_t, grid_h * grid_w, d)
    outs = []
    for i in range(0, grid_t, frame_chunk):
        part = pv[i:i + frame_chunk]
        thw = video_grid_thw.clone()
        thw[0, 0] = part.shape[0]
        outs.append(self.visual(part.reshape(-1, d), grid_thw=thw).pooler_output)
    return torch.cat(outs)


def _lm(self, cache, inputs_embeds, position_ids, **kwargs):
    """one language-model call; the 3-D ids of this call ride on the cache for the per-layer rotary"""
    cache.retake_position_ids = position_ids.clone() if isinstance(cache, PivotKVCache) else None
    cache.retake_position_ids_all = None
    if isinstance(cache, PivotKVCache) and cache.pos_embed_reforge:
        cache.retake_position_ids_all = cache.rebased_position_ids(position_ids, self.language_model.config.num_hidden_layers)
    cache.retake_rotary_emb = self.language_model.rotary_emb
    return self.language_model(input_ids=None, position_ids=position_ids, attention_mask=None, past_key_values=cache,
                               inputs_embeds=inputs_embeds, use_cache=True, **kwargs)


# -------------------------------------------------------------------------------------------- model forward
def retake_Qwen2VLModel_forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None,
                                inputs_embeds=None, use_cache=None, pixel_values=None, pixel_values_videos=None,
                                image_grid_thw=None, video_grid_thw=None, rope_deltas=None, mm_token_type_ids=None,
                                **kwargs):
    lv = getattr(self.config, "longvideo_kwargs", None)
    past_len = past_key_values.get_seq_length() if past_key_values is not None else 0
    chunk_size = None
    if lv and past_len == 0 and input_ids is not None and pixel_values_videos is not None and pixel_values is None:
        chunk_size = self.get_chunk_size(self.config, video_grid_thw)

    if chunk_size is None:
        if isinstance(past_key_values, PivotKVCache) and input_ids is not None:
            # decode after a ReTaKe prefill: positions count uncompressed tokens (reference :583-589)
            assert input_ids.shape[0] == 1
            q_len = input_ids.shape[1]
            seen = past_key_values.retake_seen_tokens
            pos = (torch.arange(q_len, device=input_ids.device) + seen).view(1, 1, -1) + self.rope_deltas.view(1, -1, 1)
            pos = pos.expand(3, -1, -1).contiguous()
            past_key_values.retake_seen_tokens = seen + q_len
            past_key_values.kvcache_compression = False
            out = _lm(self, past_key_values, self.get_input_embeddings()(input_ids), pos, **kwargs)
            return hf.Qwen2VLModelOutputWithPast(last_hidden_state=out.last_hidden_state, past_key_values=past_key_values,
                                                 rope_deltas=self.rope_deltas)
        return _ORIG["model_forward"](self, input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                                      past_key_values=past_key_values, inputs_embeds=inputs_embeds, use_cache=use_cache,
                                      pixel_values=pixel_values, pixel_values_videos=pixel_values_videos,
                                      image_grid_thw=image_grid_thw, video_grid_thw=video_grid_thw,
                                      rope_deltas=rope_deltas, mm_token_type_ids=mm_token_type_ids, **kwargs)

    # ------------------------------------------------------------------ ReTaKe prefill (reference :543-720)
    assert input_ids.shape[0] == 1, "Batch inference of long video is not supported yet!"
    if attention_mask is not None and not bool(attention_mask.all()):
        raise NotImplementedError("padded inputs are not supported (batch size 1, no padding)")
    if lv.get("kvcache_compression", False):
        kw = lv["kvcache_compression_kwargs"]
        if kw.get("dynamic_compression_ratio", False):
            input_length, max_len = input_ids.shape[1], kw["max_input_length"]
            kw["compression_ratio"] = 1 if input_length <= max_len else max_len / input_length
    cache = build_kvcache(self.config)

    position_ids, delta = mrope_position_ids(input_ids, self.config.video_token_id, video_grid_thw,
                                             self.config.vision_config.spatial_merge_size)
    self.rope_deltas = delta
    prompt_tokens = input_ids.shape[1]        # BEFORE visual compression: what generate()'s cache_position counts
    video_embeds = _video_features(self, pixel_values_videos, video_grid_thw)
    input_ids, attention_mask, video_embeds, _, position_ids, _, keypatches_mask = self.compress_video_tokens(
        input_ids=input_ids, attention_mask=attention_mask, video_embeds=video_embeds, cache_position=None,
        position_ids=position_ids, labels=None, video_grid_thw=video_grid_thw)

    inputs_embeds = self.get_input_embeddings()(input_ids)
    video_mask = input_ids == self.config.video_token_id
    if int(video_mask.sum()) != video_embeds.shape[0]:
        raise ValueError(f"Video features and video tokens do not match: tokens: {int(video_mask.sum())}, "
                         f"features {video_embeds.shape[0]}")
    inputs_embeds = inputs_embeds.masked_scatter(video_mask.unsqueeze(-1).expand_as(inputs_embeds),
                                                 video_embeds.to(inputs_embeds.device, inputs_embeds.dtype))
    if keypatches_mask is not None:
        keypatches_mask = torch.zeros_like(input_ids).bool().masked_scatter(video_mask, keypatches_mask)

    modality_segments = self.segment_input_ids(input_ids)
    compress = getattr(cache, "kvcache_compression", False)
    outputs = None
    for s, e, kind in modality_segments:
        if kind == "text":
            cache.kvcache_compression = False
            outputs = _lm(self, cache, inputs_embeds[:, s:e], position_ids[:, :, s:e], **kwargs)
        else:
            cache.kvcache_compression = compress
            for c in range(math.ceil((e - s) / chunk_size)):
                ss, ee = s + c * chunk_size, min(s + (c + 1) * chunk_size, e)
                if keypatches_mask is not None:
                    cache.keypatches_mask_chunk = keypatches_mask[0, ss:ee]
                _, pos_chunk, _, emb_chunk, prompt_length = self.forge_input_chunks(
                    ss, ee, modality_segments, None, position_ids, None, cache, inputs_embeds)
                if hasattr(cache, "before_forward"):
                    cache.before_forward(prompt_length=prompt_length)
                outputs = _lm(self, cache, emb_chunk, pos_chunk, **kwargs)
                if hasattr(cache, "after_forward"):
                    cache.after_forward()
            cache.keypatches_mask_chunk = None
            cache.kvcache_compression = False          # turned off for the trailing text and for decoding
    # decode positions are `cache_position[0] + rope_deltas` in the reference (qwen2_vl.py:583-589): both count the
    # ORIGINAL prompt, not the sequence DPSelect / MA-LLM shortened, so the first generated token sits right after the
    # largest prompt position whatever the visual ratio is
    cache.retake_seen_tokens = prompt_tokens
    return hf.Qwen2VLModelOutputWithPast(last_hidden_state=outputs.last_hidden_state, past_key_values=cache,
                                         rope_deltas=self.rope_deltas)


def install():
    """attach the ReTaKe forwards to the transformers 5.x classes (``monkeypatch.py:51-62``)"""
    if "model_forward" not in _ORIG:
        _ORIG["model_forward"] = hf.Qwen2VLModel.forward
        _ORIG["attention_forward"] = hf.Qwen2VLAttention.forward
    hf.Qwen2VLAttention.forward = retake_Qwen2VLAttention_forward
    for cls in (hf.Qwen2VLModel, hf.Qwen2VLForConditionalGeneration):
        cls.compress_video_tokens = retake_Qwen2VLForConditionalGeneration_compress_video_tokens
        cls.segment_input_ids = retake_Qwen2VLForConditionalGeneration_segment_input_ids
        cls.get_chunk_size = retake_Qwen2VLForConditionalGeneration_get_chunk_size
        cls.forge_input_chunks = retake_Qwen2VLForConditionalGeneration_forge_input_chunks
    hf.Qwen2VLModel.forward = retake_Qwen2VLModel_forward


def uninstall():
    if _ORIG:
        hf.Qwen2VLModel.forward = _ORIG["model_forward"]
        hf.Qwen2VLAttention.forward = _ORIG["attention_forward"]
