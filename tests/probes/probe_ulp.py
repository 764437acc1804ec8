import sys, os
ROOT='/root/repo'
for p in (ROOT, ROOT+'/video-retake_b200', ROOT+'/tests'): sys.path.insert(0,p)
import torch
from test_gpu_pivotkv import qkv, ref_head_scores_cuda, ulp_diff, _lc
lc=_lc()
for (H,KVH,L,D,alpha,seed) in ((16,8,2153,128,2.0,203),(16,8,2153,128,2.0,1),(16,8,2048,128,2.0,1),(28,4,2153,128,2.0,1),(16,8,2153,128,1.0,1),(8,8,2153,128,2.0,1),(2,1,2153,128,2.0,1)):
    q,k,v=qkv(H,KVH,L,D,alpha,seed=seed)
    hs=lc.pivot_head_scores(q,k); ref=ref_head_scores_cuda(q,k)
    d=ulp_diff(hs,ref)
    bad=(d>1).nonzero()
    print((H,KVH,L,D,alpha,seed),'max',int(d.max()),'n>0',int((d>0).sum()),'n>1',int((d>1).sum()), 'of', d.numel())
    for b in bad[:5]:
        i,j=int(b[0]),int(b[1]); print('   kvh',i,'key',j,'ours',float(hs[i,j]),'ref',float(ref[i,j]), hex(hs[i,j].view(torch.int16).item()&0xffff), hex(ref[i,j].view(torch.int16).item()&0xffff))
