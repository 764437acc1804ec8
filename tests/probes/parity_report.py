"""GPU probe: parity numbers for the record (profiles/r1_parity_report.json): kernels vs the reference's torch ops on the
same B200 at benchmark sizes."""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "video-retake_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch.nn.functional as F
from helpers import scene_video
from oracle import dpselect as od
from retake import longvideo_cache as lc
from retake import visual_compression as vc

out = {"dpselect": [], "pivot_scores": []}
for (T, N, C, kind) in ((1024, 256, 3584, "scene"), (1024, 256, 3584, "rand"), (512, 729, 1152, "scene"), (128, 256, 3584, "dup")):
    g = torch.Generator().manual_seed(T + N)
    if kind == "rand":
        x = torch.randn(T, N, C, generator=g)
    else:
        x = scene_video(g, T, N, C, dup_every=7 if kind == "dup" else 0)
    x = x.to(torch.bfloat16).cuda()
    sim = F.cosine_similarity(x[None, :-1], x[None, 1:], dim=-1)[0]
    want = torch.cat([torch.ones_like(sim[:1], dtype=torch.float), 1 - sim.float()], 0)
    got = vc.dpselect_distance(x)
    ent = {"T": T, "N": N, "C": C, "data": kind, "distances": got.numel(), "distance_mismatches": int((got != want).sum())}
    for r in (1.0, 0.5, 0.25):
        t = max(1, round(r * T))
        for sync in (False, True):
            idx, mask = vc.dpselect_select(want, t, sync)
            w_idx, peaks = od.dpselect_indices(want, t, sync, tie="torch")
            w_mask = peaks[w_idx][:, None].repeat(1, N) if sync else peaks.gather(0, w_idx)
            ent[f"r{r}_{'sync' if sync else 'patch'}"] = {"index_mismatches": int((idx.long() != w_idx).sum()),
                                                          "mask_mismatches": int((mask != w_mask.flatten()).sum()),
                                                          "key_patches": int(mask.sum())}
    out["dpselect"].append(ent)
    del x

def ref_hs(q, k):
    H, KVH, L, D = q.shape[1], k.shape[1], q.shape[2], q.shape[3]
    kr = k[:, :, None].expand(1, KVH, H // KVH, L, D).reshape(1, H, L, D)
    w = torch.matmul(q, kr.transpose(2, 3)) / math.sqrt(D)
    w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
    return w[0].sum(1).reshape(KVH, -1, L).mean(1)

for (H, KVH, L, D, alpha) in ((28, 4, 4096, 128, 1.0), (28, 4, 4096, 128, 3.0), (28, 4, 6272, 128, 1.0), (28, 4, 2304, 128, 2.0), (4, 2, 1024, 64, 3.0)):
    tot = bad = 0
    maxulp = 0
    idx_overlap = []
    for seed in range(4):
        g = torch.Generator().manual_seed(seed * 31 + L)
        q = (torch.randn(1, L, H, D, generator=g) * alpha).to(torch.bfloat16).cuda().transpose(1, 2)
        k = (torch.randn(1, L, KVH, D, generator=g) * alpha).to(torch.bfloat16).cuda().transpose(1, 2)
        got = lc.pivot_head_scores(q, k)
        want = ref_hs(q, k)
        d = (got.view(torch.int16).int() - want.view(torch.int16).int()).abs()
        tot += d.numel(); bad += int((d > 0).sum()); maxulp = max(maxulp, int(d.max()))
        keep = max(1, int(0.25 * L))
        a = lc.pivot_select(got, keep).long()
        b = want.mean(0).topk(keep).indices.sort().values
        idx_overlap.append(len(set(a.tolist()) & set(b.tolist())) / keep)
    out["pivot_scores"].append({"H": H, "KVH": KVH, "L": L, "D": D, "alpha": alpha, "head_scores": tot,
                                "differ_from_torch_cuda": bad, "fraction": bad / tot, "max_ulp": maxulp,
                                "kept_index_overlap_r0.25": idx_overlap})
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
