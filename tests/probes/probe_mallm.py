"""GPU probe for the MA-LLM path: which ATen-CUDA rules do torch.max(dim) ties and bf16 mean(-1) follow, and does the
explicit oracle (oracle/mallm.py, reduce="aten_cuda") reproduce the stock torch op sequence executed on the GPU?"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from oracle import mallm, reference_ops as ro
from oracle.dpselect import _r
from helpers import scene_video

dev = torch.device("cuda")
out = {}
# Q1: torch.max(dim=1) on [1, T, N] bf16 with ties / NaN
g = torch.Generator().manual_seed(1)
x = (torch.randint(0, 4, (1, 300, 97), generator=g).float() / 4).to(torch.bfloat16)
x[0, 17, 5] = float("nan"); x[0, 200, 5] = float("nan"); x[0, 40, 6] = float("inf")
got = torch.max(x.to(dev), dim=1, keepdim=True)[1][0, 0].cpu()
want = mallm.first_argmax(x[0].float())
out["Q1_max_ties_lowest_index_nan_first"] = bool(torch.equal(got, want))
out["Q1_detail"] = {"got5": int(got[5]), "want5": int(want[5]), "n_diff": int((got != want).sum())}
# Q2: bf16 mean(-1) order
res = {}
for n in (8, 24, 64, 96, 127, 128, 129, 136, 200, 256, 729, 1000):
    v = torch.randn(37, n, generator=g).to(torch.bfloat16)
    gpu = v.to(dev).mean(-1).float().cpu()
    row = {}
    for vec in (4, 8):
        mine = mallm.aten_cuda_bf16_rowmean(v.float(), vec)
        row[f"vec{vec}_mismatch"] = int((mine != gpu).sum())
    # alignment-free variant: the tensor inside the reference is a fresh contiguous [1, T-1, N] -> row r starts at r*n*2
    res[n] = row
out["Q2_bf16_mean"] = res
# Q3: whole rounds, oracle vs stock ops on the GPU
q3 = {}
for (T, N, C, t) in ((24, 96, 256, 9), (20, 136, 512, 11), (16, 256, 1152, 8), (12, 729, 256, 6)):
    for sync in (False, True):
        for hard in (False, True):
            x = scene_video(g, T, N, C, dup_every=5).to(torch.bfloat16)[None]
            want, wsz = ro.mallm_compress(x.to(dev), t, sync, hard)
            best = None
            for vec in (4, 8):
                mine, msz = mallm.mallm_compress(x, t, sync, hard, reduce="aten_cuda", mean_vec=vec)
                ok = bool(torch.equal(mine, want.cpu())) and (hard or bool(torch.equal(msz, wsz.cpu())))
                q3[f"T{T}_N{N}_C{C}_t{t}_sync{int(sync)}_hard{int(hard)}_vec{vec}"] = ok
out["Q3_rounds"] = q3
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_mallm.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
