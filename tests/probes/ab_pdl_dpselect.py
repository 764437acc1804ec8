"""GPU probe: whole DPSelect operator (dis + select + gather) timed with CUDA events, run twice by the caller with and
without RTK_NO_PDL=1."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "video-retake_b200")):
    sys.path.insert(0, p)
import torch
from retake import visual_compression as vc
T, N, C = 1024, 256, 3584
x = torch.randn(T, N, C, device="cuda").to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {}
def t(fn, n=12):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
res["operator_t1024"] = t(lambda: vc.memory_bank_compress_keyframe(x[None], T, 3, sync=False))
res["operator_t512"] = t(lambda: vc.memory_bank_compress_keyframe(x[None], T // 2, 3, sync=False))
res["dis"] = t(lambda: vc.dpselect_distance(x))
dis = vc.dpselect_distance(x)
idx, _ = vc.dpselect_select(dis, T, False)
res["select"] = t(lambda: vc.dpselect_select(dis, T, False))
res["gather"] = t(lambda: vc.dpselect_gather(x, idx, False))
print(os.environ.get("RTK_NO_PDL", "0"), json.dumps(res))
