"""Drop-in for ``retake/monkeypatch.py`` of SCZwangxiao/video-ReTaKe: the same four public functions.

``patch_*_config`` attach ``longvideo_kwargs`` (and the YaRN settings) exactly like the reference
(``monkeypatch.py:24-48``).  ``patch_qwen2vl`` / ``patch_llava_onevision`` install the model-level glue
(``retake/qwen2_vl.py`` / ``retake/llava_onevision.py`` in this package, re-targeted at transformers 5.x) that
calls the two B200 operators; any ``method`` other than ``"retake"`` raises ``NotImplementedError`` like the
reference (``monkeypatch.py:63-64,78-79``).
"""
from __future__ import annotations


def _rope_dict(config):
    """transformers 5.x keeps the rope settings in ``rope_parameters``; 4.48 used ``rope_scaling``."""
    for name in ("rope_parameters", "rope_scaling"):
        d = getattr(config, name, None)
        if isinstance(d, dict):
            return d
    d = {}
    config.rope_scaling = d
    return d


def patch_qwen2vl_config(config, exp_configs):
    if "scaling_factor" in exp_configs:
        for cfg in {id(config): config, id(getattr(config, "text_config", config)): getattr(config, "text_config", config)}.values():
            rope = _rope_dict(cfg)
            rope.pop("type", None)
            rope["rope_type"] = "yarn"
            rope["factor"] = exp_configs["scaling_factor"]
            rope["beta_fast"] = 32.0
            rope["beta_slow"] = 1.0
    config.longvideo_kwargs = exp_configs.get("longvideo_kwargs", {})
    return config


def patch_llava_onevision_config(config, exp_configs):
    if "scaling_factor" in exp_configs:
        rope = {"rope_type": "yarn", "factor": exp_configs["scaling_factor"], "beta_fast": 32.0, "beta_slow": 1.0}
        old = getattr(config.text_config, "rope_parameters", None)
        if isinstance(old, dict):
            if "rope_theta" in old:
                rope["rope_theta"] = old["rope_theta"]
            config.text_config.rope_parameters = rope
        else:
            config.text_config.rope_scaling = rope
    config.longvideo_kwargs = exp_configs.get("longvideo_kwargs", {})
    return config


def patch_qwen2vl(method):
    if method == "retake":
        from . import qwen2_vl
        print("Using ReTaKe for Qwen2VLForConditionalGeneration!")
        qwen2_vl.install()
    else:
        raise NotImplementedError


def patch_llava_onevision(method):
    if method == "retake":
        from . import llava_onevision
        print("Using ReTaKe for LlavaOnevisionForConditionalGeneration!")
        llava_onevision.install()
    else:
        raise NotImplementedError
