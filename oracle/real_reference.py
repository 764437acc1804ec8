"""Loader for the UNMODIFIED reference modules of the hot path.  TEST / BASELINE INFRASTRUCTURE ONLY.

Looks for a directory that contains ``retake/visual_compression.py`` and ``retake/longvideo_cache.py`` of
SCZwangxiao/video-ReTaKe in this order: ``$RETAKE_REFERENCE``, ``baseline/_ref`` (git-ignored copy made by
``__graft_entry__.build()`` when ``/root/reference`` is present - it travels to the GPU box with the snapshot, the
reference tree itself does not), ``/root/reference``.  The two modules import and run unmodified on torch 2.11 /
transformers 5.5; ``PivotKVCache.update`` assigns ``self.key_cache[i]`` (transformers 4.48 naming,
``longvideo_cache.py:313``), so the class is instantiated through a subclass that only adds list views onto
``layers[i].keys / .values`` - the same 10-line shim ``tests/golden/make_golden.py`` pinned the fixtures with.
``/root/reference`` is only looked at when the caller asks for it (``allow_system_tree=True``: CPU tests in the build
container); ``bench.py`` and the GPU tests never read it - they use the copy that travelled with the snapshot.

Used by ``bench.py`` (the ``cpu_baseline`` leg and ``--impl reference``: ``kind: "reference"``) and by tests; nothing under
``video-retake_b200/`` imports it."""
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.environ.get("RETAKE_REFERENCE"), os.path.join(ROOT, "baseline", "_ref")]
SYSTEM_TREE = "/root/reference"
FILES = ("visual_compression.py", "longvideo_cache.py")


def find(allow_system_tree=False):
    for c in CANDIDATES + ([SYSTEM_TREE] if allow_system_tree else []):
        if c and all(os.path.isfile(os.path.join(c, "retake", f)) for f in FILES):
            return c
    return None


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


class _LayerListView:
    def __init__(self, cache, attr):
        self._c, self._a = cache, attr

    def __getitem__(self, i):
        return getattr(self._c.layers[i], self._a)

    def __setitem__(self, i, v):
        setattr(self._c.layers[i], self._a, v)

    def __len__(self):
        return len(self._c.layers)


_LOADED = None


def load(allow_system_tree=False):
    """-> (visual_compression module, shimmed PivotKVCache class, directory) or None when no reference tree is around"""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    base = find(allow_system_tree)
    if base is None:
        return None
    vc = _load(os.path.join(base, "retake", "visual_compression.py"), "retake_reference_visual_compression")
    lc = _load(os.path.join(base, "retake", "longvideo_cache.py"), "retake_reference_longvideo_cache")

    class ShimPivotKVCache(lc.PivotKVCache):
        @property
        def key_cache(self):
            return _LayerListView(self, "keys")

        @property
        def value_cache(self):
            return _LayerListView(self, "values")

    _LOADED = (vc, ShimPivotKVCache, base)
    return _LOADED


def llm_config(heads, kv_heads, head_dim, layers, ratio, reforge):
    cfg = types.SimpleNamespace(hidden_size=heads * head_dim, num_hidden_layers=layers, num_attention_heads=heads,
                                num_key_value_heads=kv_heads)
    cfg.longvideo_kwargs = {"kvcache_compression": True,
                            "kvcache_compression_kwargs": {"compression_ratio": ratio, "compression_method": "pivotkv",
                                                           "pos_embed_reforge": reforge}}
    return cfg
