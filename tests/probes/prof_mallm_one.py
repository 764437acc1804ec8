"""GPU probe: one fused MA-LLM call per configuration (for an ncu launch list)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "video-retake_b200")):
    sys.path.insert(0, p)
import torch
from retake import visual_compression as vc
T, N, C, t = 512, 256, 3584, 256
x = torch.randn(1, T, N, C, device="cuda").to(torch.bfloat16)
for sync in (False, True):
    for hard in (False, True):
        vc.mallm_compress(x, t, sync=sync, hard=hard)
torch.cuda.synchronize()
