"""GPU probe: DPSelect kernels at the 2048-frame Qwen shape (or LLaVA shape with SHAPE=llava) for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "video-retake_b200")):
    sys.path.insert(0, p)
import torch
from retake import visual_compression as vc
T, N, C = (1024, 256, 3584) if os.environ.get("SHAPE", "qwen") == "qwen" else (1024, 729, 1152)
x = torch.randn(T, N, C, device="cuda").to(torch.bfloat16)
for _ in range(3):
    dis = vc.dpselect_distance(x)
    idx, mask = vc.dpselect_select(dis, T // 2, False)
    out = vc.dpselect_gather(x, idx, False)
torch.cuda.synchronize()
print(float(dis.mean()), int(mask.sum()))
