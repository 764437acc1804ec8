"""2-GPU probe (torchrun): does torch symmetric memory give peer-mapped pointers on this box, what does a 32 KB exchange
cost through it, and what does the same exchange cost as ONE NCCL all_gather_into_tensor?  Prints one JSON line on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
out = {"world": world}
try:
    import torch.distributed._symmetric_memory as sm
    t = sm.empty((4, 8192), dtype=torch.bfloat16, device=dev)
    h = sm.rendezvous(t, dist.group.WORLD)
    out["symm"] = {"buffer_ptrs": [hex(p) for p in h.buffer_ptrs], "signal_pad_ptrs": [hex(p) for p in h.signal_pad_ptrs],
                   "signal_pad_size": h.signal_pad_size, "multicast": bool(h.has_multicast_support)}
    t.fill_(float(rank + 1))
    dist.barrier()
    torch.cuda.synchronize()
    peer = (rank + 1) % world
    pb = h.get_buffer(peer, (4, 8192), torch.bfloat16)
    out["symm"]["peer_read_ok"] = bool((pb == float(peer + 1)).all())
    # time: write 16 KB into the peer + a flag, 50 times
    mine = torch.full((2, 8192), float(rank), dtype=torch.bfloat16, device=dev)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    a.record()
    for _ in range(50):
        pb[2 * rank:2 * rank + 2].copy_(mine)
    b.record()
    torch.cuda.synchronize()
    out["symm"]["p2p_copy_16KB_us"] = a.elapsed_time(b) / 50 * 1e3
except Exception as e:  # noqa: BLE001
    out["symm_error"] = repr(e)[:500]
try:
    loc = torch.full((2, 8192), float(rank), dtype=torch.bfloat16, device=dev)
    full = torch.empty((2 * world, 8192), dtype=torch.bfloat16, device=dev)
    for _ in range(5):
        dist.all_gather_into_tensor(full, loc)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        dist.all_gather_into_tensor(full, loc)
    b.record()
    torch.cuda.synchronize()
    out["nccl_all_gather_into_tensor_32KB_us"] = a.elapsed_time(b) / 50 * 1e3
    parts = [torch.empty_like(loc) for _ in range(world)]
    a.record()
    for _ in range(50):
        dist.all_gather(parts, loc)
        torch.cat(parts)
    b.record()
    torch.cuda.synchronize()
    out["nccl_list_all_gather_plus_cat_us"] = a.elapsed_time(b) / 50 * 1e3
except Exception as e:  # noqa: BLE001
    out["nccl_error"] = repr(e)[:300]
# CUDA IPC through torch storages (fallback transport)
try:
    buf = torch.zeros(4 * 8192, dtype=torch.bfloat16, device=dev)
    handle = buf.untyped_storage()._share_cuda_()
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    peer = (rank + 1) % world
    st = torch.UntypedStorage._new_shared_cuda(*handles[peer])
    pt = torch.empty(0, dtype=torch.bfloat16, device=st.device).set_(st)
    out["ipc"] = {"peer_device": str(pt.device), "numel": pt.numel()}
    buf.fill_(float(rank + 7))
    dist.barrier()
    torch.cuda.synchronize()
    out["ipc"]["peer_read_ok"] = bool((pt.to(dev) == float(peer + 7)).all())
except Exception as e:  # noqa: BLE001
    out["ipc_error"] = repr(e)[:500]
if rank == 0:
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
