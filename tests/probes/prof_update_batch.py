"""GPU probe: N chunks of 28 layers through the deferred (batched) PivotKV path at the benchmark shape (for ncu:
ncu --set full --clock-control none --import-source on -k regex:pivot_score_kernel --launch-skip 2 --launch-count 2)."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "video-retake_b200")):
    sys.path.insert(0, p)
import torch
import bench
from retake import longvideo_cache as lc
H, KVH, D, L, LAYERS = 28, 4, 128, int(os.environ.get("L", "4096")), int(os.environ.get("LAYERS", "28"))
dev = torch.device("cuda")
s = types.SimpleNamespace(H=H, KVH=KVH, D=D, layers=LAYERS, kv_ratio=500 / 4096, reforge=True, deferred=True)
cache = lc.PivotKVCache(bench.cache_config(s))
rot = bench.make_rotary(dev)
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(LAYERS, L, H, D, device=dev, generator=g).to(torch.bfloat16)
k = torch.randn(LAYERS, L, KVH, D, device=dev, generator=g).to(torch.bfloat16)
v = torch.randn(LAYERS, L, KVH, D, device=dev, generator=g).to(torch.bfloat16)
ar = torch.arange(L, device=dev)
pos_grid = torch.stack([ar // 256, (ar % 256) // 16, ar % 16])[:, None]
for c in range(int(os.environ.get("N", "2"))):
    cache.kvcache_compression = True
    cache.keypatches_mask_chunk = (torch.rand(L, device=dev, generator=g) < 0.15)
    for layer in range(LAYERS):
        pos = pos_grid.clone()
        pos[0] += cache.get_prev_temporal_idx(layer) + 1
        cache.update(k[layer:layer + 1].transpose(1, 2), v[layer:layer + 1].transpose(1, 2), layer,
                     {"query_states": q[layer:layer + 1].transpose(1, 2), "position_ids": pos, "rotary_emb": rot,
                      "mrope_section": [16, 24, 24]})
    cache.after_forward()
torch.cuda.synchronize()
print(cache.get_seq_length(0), cache.last_keep_indices[:4].tolist())
