"""GPU probe: CUDA-event breakdown of one DPSelect call and one PivotKV update at the benchmark shapes, plus
host enqueue time per update.  Not part of the test-suite; writes gpurun_out/breakdown.json."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "video-retake_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from retake import longvideo_cache as lc  # noqa: E402
from retake import visual_compression as vc  # noqa: E402
import bench  # noqa: E402

dev = torch.device("cuda", 0)
out = {}


def timeit(fn, n=10, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return {"median_ms": ts[len(ts) // 2], "min_ms": ts[0]}


flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

# ------------------------------------------------------------------ DPSelect at T = 1024 (2048 frames) and LLaVA shape
for name, (T, N, C) in {"qwen_2048f": (1024, 256, 3584), "qwen_256f": (128, 256, 3584), "llava_512f": (512, 729, 1152)}.items():
    x = torch.randn(T, N, C, device=dev).to(torch.bfloat16)
    ent = {}
    r = timeit(lambda: vc.dpselect_distance(x), flush=flush)
    by = 2.0 * T * N * C + 4.0 * T * N
    ent["dis"] = {**r, "GBps": by / r["median_ms"] / 1e6}
    dis = vc.dpselect_distance(x)
    for t in (T, T // 2):
        r = timeit(lambda: vc.dpselect_select(dis, t, False))
        ent[f"select_t{t}"] = r
        idx, _ = vc.dpselect_select(dis, t, False)
        r = timeit(lambda: vc.dpselect_gather(x, idx, False), flush=flush)
        ent[f"gather_t{t}"] = {**r, "GBps": 4.0 * t * N * C / r["median_ms"] / 1e6}
    r = timeit(lambda: vc.dpselect_select(dis, T // 2, True))
    ent["select_sync"] = r
    out[name] = ent
    del x

# ------------------------------------------------------------------ one PivotKV update, L = 4096, 7B shape
H, KVH, D, L = 28, 4, 128, 4096
q = torch.randn(1, L, H, D, device=dev).to(torch.bfloat16).transpose(1, 2)
k = torch.randn(1, L, KVH, D, device=dev).to(torch.bfloat16).transpose(1, 2)
v = torch.randn(1, L, KVH, D, device=dev).to(torch.bfloat16).transpose(1, 2)
rot = bench.make_rotary(dev)
ar = torch.arange(L, device=dev)
pos = torch.stack([ar // 256, (ar % 256) // 16, ar % 16])[:, None]
mask = torch.rand(L, device=dev) < 0.15
cos, sin = rot(v, pos)
pk = {}
pk["rotary_emb_call"] = timeit(lambda: rot(v, pos))
pk["rope_q"] = timeit(lambda: lc.pivot_rope(q, cos, sin, [16, 24, 24], rot.attention_scaling))
pk["rope_k"] = timeit(lambda: lc.pivot_rope(k, cos, sin, [16, 24, 24], rot.attention_scaling))
qu = lc.pivot_rope(q, cos, sin, [16, 24, 24], rot.attention_scaling)
ku = lc.pivot_rope(k, cos, sin, [16, 24, 24], rot.attention_scaling)
r = timeit(lambda: lc.pivot_head_scores(qu, ku), n=20)
pk["score"] = {**r, "TFLOPs_algorithmic": 2.0 * H * L * L * D / r["median_ms"] / 1e9}
r = timeit(lambda: lc.pivot_head_scores(q, k), n=20)
pk["score_strided_inputs"] = {**r, "TFLOPs_algorithmic": 2.0 * H * L * L * D / r["median_ms"] / 1e9}
hs = lc.pivot_head_scores(qu, ku)
for keep in (500, 1024):
    pk[f"select_keep{keep}"] = timeit(lambda: lc.pivot_select(hs, keep, mask))
    idx = lc.pivot_select(hs, keep, mask)
    r = timeit(lambda: lc.pivot_compact(ku, v, idx, pos, True))
    pk[f"compact_keep{keep}"] = {**r, "GBps": 8192.0 * keep / r["median_ms"] / 1e6}
past_k = torch.randn(1, KVH, 16000, D, device=dev).to(torch.bfloat16)
pk["cat_past16k_chunk"] = timeit(lambda: torch.cat([past_k, k], dim=-2))

# whole update through the public API: device time and host enqueue time
cfg = bench.cache_config(type("S", (), {"H": H, "D": D, "layers": 1, "KVH": KVH, "kv_ratio": 0.122, "reforge": True, "deferred": False})())


def one_update(cache):
    cache.keypatches_mask_chunk = mask
    cache.update(k, v, 0, {"query_states": q, "position_ids": pos.clone(), "rotary_emb": rot, "mrope_section": [16, 24, 24]})


def fresh_updates(n):
    cache = lc.build_kvcache(cfg)
    for _ in range(n):
        one_update(cache)


pk["update_x8_public_api"] = timeit(lambda: fresh_updates(8), n=5)
torch.cuda.synchronize()
cache = lc.build_kvcache(cfg)
one_update(cache)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(16):
    one_update(cache)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
pk["host_enqueue_ms_per_update"] = (t1 - t0) / 16 * 1e3
pk["wall_ms_per_update"] = (t2 - t0) / 16 * 1e3
out["pivotkv_L4096"] = pk

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "breakdown.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
