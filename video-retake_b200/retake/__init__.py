"""B200-native DPSelect + PivotKV behind the module layout of SCZwangxiao/video-ReTaKe (``retake/``).

Put ``video-retake_b200/`` on ``PYTHONPATH`` (the reference is used with ``PYTHONPATH=./``) and
``from retake.monkeypatch import patch_qwen2vl`` etc. resolve to this package.
"""
