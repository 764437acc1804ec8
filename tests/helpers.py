"""Shared synthetic-input helpers for tests, golden generation and bench (no reference, no oracle)."""
import torch


class TableRotary:
    """Deterministic stand-in for the HF rotary module: callable (x, position_ids) -> (cos, sin)
    with ``attention_scaling`` (``longvideo_cache.py:249,256``).  YaRN-like scaling 1.1386."""

    def __init__(self, head_dim, base=10000.0, attention_scaling=1.1386, mrope=True):
        self.inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
        self.attention_scaling = attention_scaling
        self.mrope = mrope

    def __call__(self, x, position_ids):
        pos = position_ids.to(torch.float32)             # [3,1,L] or [1,L]
        freqs = pos[..., None] * self.inv_freq           # [...,L,D/2]
        emb = torch.cat((freqs, freqs), dim=-1)
        cos = emb.cos() * self.attention_scaling
        sin = emb.sin() * self.attention_scaling
        return cos.to(x.dtype), sin.to(x.dtype)


def scene_video(g, T, N, C, sigma=0.3, dup_every=0):
    """scene-structured embeddings: x[t] = scene[s(t)] + sigma * eps[t]."""
    x = torch.empty(T, N, C)
    t = 0
    while t < T:
        run = int(torch.randint(2, 7, (1,), generator=g))
        scene = torch.randn(N, C, generator=g)
        for _ in range(run):
            if t >= T:
                break
            x[t] = scene + sigma * torch.randn(N, C, generator=g)
            t += 1
    if dup_every:
        for t in range(dup_every, T, dup_every):
            x[t] = x[t - 1]                               # exact duplicates -> dis == 0 ties
    return x
