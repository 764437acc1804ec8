"""GPU probe: a few rtk_pivot_score calls at the benchmark shape (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "video-retake_b200")):
    sys.path.insert(0, p)
import torch
from retake import longvideo_cache as lc
H, KVH, D = 28, 4, 128
L = int(os.environ.get("L", "4096"))
q = torch.randn(1, L, H, D, device="cuda").to(torch.bfloat16).transpose(1, 2)
k = torch.randn(1, L, KVH, D, device="cuda").to(torch.bfloat16).transpose(1, 2)
for _ in range(int(os.environ.get("N", "4"))):
    hs = lc.pivot_head_scores(q, k)
torch.cuda.synchronize()
print(float(hs.float().mean()))
