"""GPU probe (torchrun, N ranks): one video split across the GPUs of a box - PivotKV by KV head, DPSelect by frame range
(retake/distributed.py) - timed against the single-GPU operators on rank 0.  Writes gpurun_out/sharded_N.json."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "video-retake_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist
from helpers import TableRotary
from retake import distributed as rd
from retake import longvideo_cache as lc
from retake import visual_compression as vc

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator().manual_seed(0)
H, KVH, L, D, mrope, keep = 28, 4, 4096, 128, [16, 24, 24], 1024
q = torch.randn(1, L, H, D, generator=g).to(torch.bfloat16).to(dev).transpose(1, 2)
k = torch.randn(1, L, KVH, D, generator=g).to(torch.bfloat16).to(dev).transpose(1, 2)
v = torch.randn(1, L, KVH, D, generator=g).to(torch.bfloat16).to(dev).transpose(1, 2)
ar = torch.arange(L, device=dev)
pos = torch.stack([ar // 256, (ar % 256) // 16, ar % 16])[:, None]
rot = TableRotary(D)
rot.inv_freq = rot.inv_freq.to(dev)
per = [KVH // world] * world
g0, G = rank * per[0], H // KVH
ql, kl, vl = q[:, g0 * G:(g0 + per[0]) * G], k[:, g0:g0 + per[0]], v[:, g0:g0 + per[0]]

def timed(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

out = {"world": world}
for reforge in (False, True):
    out[f"pivot_sharded_reforge{int(reforge)}_ms"] = timed(lambda: rd.pivot_update_kv_sharded(
        ql, kl, vl, keep, per, None, pos, rot, mrope, reforge))
    if rank == 0:
        def single():
            hs = lc.pivot_head_scores(q, k)
            idx = lc.pivot_select(hs, keep, None)
            lc.pivot_compact(k, v, idx, pos, reforge=False)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(5): single()
        torch.cuda.synchronize(); a.record()
        for _ in range(20): single()
        b.record(); torch.cuda.synchronize()
        out["pivot_single_gpu_unfused_ms"] = a.elapsed_time(b) / 20
    dist.barrier()
T, N, C = 1024, 256, 3584
x = torch.randn(T, N, C, generator=g).to(torch.bfloat16).to(dev)
t0, t1 = rd.split_range(T, world)[rank]
xl = x[t0 - int(t0 > 0):t1].contiguous()
out["dpselect_sharded_ms"] = timed(lambda: rd.dpselect_frame_sharded(xl, t0, t1, T, T // 2, False))
if rank == 0:
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): vc.memory_bank_compress_keyframe(x[None], T // 2, 3, sync=False)
    torch.cuda.synchronize(); a.record()
    for _ in range(10): vc.memory_bank_compress_keyframe(x[None], T // 2, 3, sync=False)
    b.record(); torch.cuda.synchronize()
    out["dpselect_single_gpu_ms"] = a.elapsed_time(b) / 10
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"sharded_{world}.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
dist.barrier()
dist.destroy_process_group()
