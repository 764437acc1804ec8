"""GPU probe: a head-score entry that differs by 2 bf16 steps from the cuBLAS path - which logits cause it?"""
import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, ROOT + '/video-retake_b200', ROOT + '/tests'): sys.path.insert(0, p)
import torch
from test_gpu_pivotkv import qkv, ref_head_scores_cuda, _lc
lc = _lc()
H, KVH, L, D = 16, 8, 2153, 128
q, k, v = qkv(H, KVH, L, D, 2.0, seed=203)
hs = lc.pivot_head_scores(q, k); ref = ref_head_scores_cuda(q, k)
BF = torch.bfloat16
for (kv, key) in ((7, 46), (7, 427)):
    print("entry", kv, key, "ours", float(hs[kv, key]), "cublas-ref", float(ref[kv, key]))
    for h in (2 * kv, 2 * kv + 1):
        qh = q[0, h]                      # [L, D]
        kh = k[0, kv]                     # [L, D]
        S_cublas = (qh @ kh.t())          # bf16 out, fp32 accumulate (cuBLAS)
        S_exact = (qh.double() @ kh.double().t()).to(BF)     # exact dot products, one rounding
        def colsum(S):
            w = S / math.sqrt(D)
            p = torch.softmax(w, dim=-1, dtype=torch.float32).to(BF)
            return p[:, key].float().sum(), p[:, key]
        c1, p1 = colsum(S_cublas); c2, p2 = colsum(S_exact)
        nflip = int((S_cublas != S_exact).sum()); 
        print(f"  head {h}: colsum cublas {float(c1):.6f} exact-logits {float(c2):.6f}; logits differing cublas vs exact: {nflip} of {S_exact.numel()}; "
              f"in this column: {int((S_cublas[:, key] != S_exact[:, key]).sum())}; max p in column {float(p2.float().max()):.4f}; rows where p differs {int((p1 != p2).sum())}")
        rows = (p1 != p2).nonzero()[:, 0][:4]
        for r in rows:
            r = int(r)
            print(f"     query {r}: p cublas {float(p1[r]):.6f} exact {float(p2[r]):.6f}; row logits differing {int((S_cublas[r] != S_exact[r]).sum())}; max logit {float(S_exact[r].float().max()):.3f}")
