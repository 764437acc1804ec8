"""CPU: the C-ABI library loads and exports every symbol include/rtk_b200.h declares; argument validation
paths that need no GPU; the host wrappers refuse CPU tensors (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "rtk_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rtk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from retake import _native
    lib = _native.lib()
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in rtk_b200.h but not exported"
    assert sorted(_native.EXPORTS) == names
    assert lib.rtk_version() == _native.ABI_VERSION == 3
    assert lib.rtk_error_string(0) == b"ok"
    assert b"aligned" in lib.rtk_error_string(-2)


def test_argument_errors_without_gpu():
    from retake import _native
    lib = _native.lib()
    assert lib.rtk_dpselect_dis(None, 4, 4, 256, 0, None, None) == -1          # RTK_E_BADARG
    buf = ctypes.create_string_buffer(64)
    a = ctypes.addressof(buf)
    a16 = (a + 15) & ~15
    assert lib.rtk_dpselect_dis(a16, 4, 4, 100, 0, a16, None) == -3            # C % 8 / C < 256: RTK_E_UNSUPPORTED
    assert lib.rtk_dpselect_dis(a16 + 2, 4, 4, 256, 0, a16, None) == -2        # RTK_E_ALIGN
    assert lib.rtk_dpselect_select(a16, 4, 4, 9, 0, a16, a16, None) == -1      # t > T
    assert lib.rtk_pivot_score(a16, 4, 128, 512, a16, 3, 128, 384, 16, 128, a16, a16, 1 << 20, None) == -3  # H % KVH
    assert lib.rtk_pivot_score(a16, 4, 128, 512, a16, 2, 128, 256, 16, 96, a16, a16, 1 << 20, None) == -3   # D
    assert lib.rtk_pivot_score(a16, 4, 128, 512, a16, 2, 128, 256, 16, 128, a16, a16, 8, None) == -4        # workspace
    assert lib.rtk_pivot_score_workspace_bytes(28, 4096) == 10 * 28 * 4096 * 4
    assert lib.rtk_pivot_score_workspace_bytes(4, 130) == 10 * 4 * 256 * 4
    assert lib.rtk_pivot_select(a16, 4, 16, None, 17, a16, None, None) == -1   # keep > L
    # batched update: workspace query, argument errors
    one = lib.rtk_pivot_update_batch_workspace_bytes(28, 4, 4096, 128, 1)
    assert lib.rtk_pivot_update_batch_workspace_bytes(28, 4, 4096, 128, 28) > 27 * (one - 1024) > 0
    assert lib.rtk_pivot_update_batch_workspace_bytes(28, 4, 4096, 128, 64) == lib.rtk_pivot_update_batch_workspace_bytes(28, 4, 4096, 128, 32)
    assert lib.rtk_pivot_update_batch(None, 2, a16, 1 << 20, None) == -1
    from retake.longvideo_cache import _UpdateArgs
    arr = (_UpdateArgs * 2)()
    a256 = (ctypes.addressof(ctypes.create_string_buffer(1024)) + 255) & ~255
    assert lib.rtk_pivot_update_batch(arr, 2, a256, 512, None) == -1           # H == 0
    for x in arr:
        x.H, x.KVH, x.L, x.D, x.keep = 4, 2, 64, 64, 16
    assert lib.rtk_pivot_update_batch(arr, 2, a256, 512, None) == -4           # workspace too small
    arr[0].reforge = 1
    assert lib.rtk_pivot_update_batch(arr, 2, a256, 1 << 30, None) == -3       # reforge without inv_freq: not batchable


def test_host_wrappers_refuse_cpu_tensors():
    from retake import _native
    from retake.visual_compression import memory_bank_compress_keyframe
    x = torch.zeros(1, 4, 4, 256, dtype=torch.bfloat16)
    with pytest.raises(_native.RtkError):
        memory_bank_compress_keyframe(x, 2, 3, sync=False)


def test_build_kvcache_factory():
    import types
    from transformers.cache_utils import DynamicCache
    from retake.longvideo_cache import PivotKVCache, build_kvcache
    cfg = types.SimpleNamespace(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2)
    assert type(build_kvcache(cfg)) is DynamicCache
    cfg.longvideo_kwargs = {"kvcache_compression": False}
    assert type(build_kvcache(cfg)) is DynamicCache
    cfg.longvideo_kwargs = {"kvcache_compression": True,
                            "kvcache_compression_kwargs": {"compression_ratio": 0.5, "compression_method": "PivotKV",
                                                           "pos_embed_reforge": True}}
    c = build_kvcache(cfg)
    assert isinstance(c, PivotKVCache) and isinstance(c, DynamicCache)
    assert c.head_dim == 64 and c.num_key_value_groups == 2 and c.pos_embed_reforge and c.kvcache_compression
    assert c.get_prev_temporal_idx(0) == -1 and c.num_evicted_tokens == [] and c.position_cache == []
    assert c.deferred_compression is False                 # default: compression inside update(), the reference's call order
    cfg.longvideo_kwargs["kvcache_compression_kwargs"]["deferred_compression"] = True
    d = build_kvcache(cfg)
    assert d.deferred_compression is True and d._deferred == []
    d.after_forward()                                      # nothing pending: no library call, works without a GPU
    del cfg.longvideo_kwargs["kvcache_compression_kwargs"]["deferred_compression"]
    cfg.longvideo_kwargs["kvcache_compression_kwargs"]["compression_method"] = "h2o"
    with pytest.raises(NotImplementedError):
        build_kvcache(cfg)
