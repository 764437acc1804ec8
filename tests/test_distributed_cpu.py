"""world_size-2 gloo tests (CPU) of the host logic that splits ONE video across GPUs (retake/distributed.py):
frame-range partition with a one-frame halo, uneven all-gather, ownership of output slots, reassembly, and the
KV-head score gather.  The CUDA kernels are replaced by the oracle here; the 2-GPU NCCL run is in test_gpu_distributed.py."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, sync):
    for p in (ROOT, os.path.join(ROOT, "video-retake_b200"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    from oracle import dpselect as od
    from oracle import pivotkv as op
    from retake import distributed as rd
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        # ---- uneven all-gather
        sizes = [3, 5][:world]
        loc = torch.full((sizes[rank], 2), float(rank))
        full = rd.all_gather_rows(loc, sizes)
        assert full.shape == (sum(sizes), 2) and bool((full[:3] == 0).all()) and bool((full[3:] == 1).all())
        assert rd.split_range(19, 2) == [(0, 10), (10, 19)] and rd.split_range(4, 4) == [(0, 1), (1, 2), (2, 3), (3, 4)]

        # ---- DPSelect split by frame range, oracle standing in for the kernels
        g = torch.Generator().manual_seed(5)
        T, N, C, t = 19, 6, 32, 8
        x = torch.randn(T, N, C, generator=g)
        x[7] = x[6]                                             # exact duplicate across ... (and 10 | 9 is the seam)
        x[10] = x[9]
        t0, t1 = rd.split_range(T, world)[rank]
        halo = int(t0 > 0)
        dis_local = od.adjacent_cosine_distance(x[t0 - halo:t1])[halo:]          # rank 0 keeps its row of ones
        if halo:
            assert dis_local.shape[0] == t1 - t0
        dis = rd.all_gather_rows(dis_local, [b - a for a, b in rd.split_range(T, world)])
        want_dis = od.adjacent_cosine_distance(x)
        assert torch.equal(dis, want_dis), "halo split must reproduce the whole distance matrix bit for bit"
        idx, peaks = od.dpselect_indices(dis, t, sync)
        slots, frames = rd.owned_slots(idx, t0, t1, N)
        assert bool(((frames >= t0) & (frames < t1)).all())
        rows = x[frames, slots % N]                              # what rtk_gather_rows does on the GPU
        out = rd.assemble_compacted(rows, slots, t, N)
        want_out, want_mask = od.memory_bank_compress_keyframe(x[None], t, 3, sync=sync)
        assert torch.equal(out, want_out)
        # the sync-free form writes owned rows straight into their slots of a zero-filled [1, t, N, C] buffer: the sum of the
        # ranks' bit patterns is the whole tensor (what rtk_dpselect_gather_owned + assemble_owned do on the GPU)
        full_idx = idx if idx.dim() == 2 else idx[:, None].expand(-1, N)
        own = (full_idx >= t0) & (full_idx < t1)
        part = torch.zeros(1, t, N, C, dtype=torch.bfloat16)
        xb = x.to(torch.bfloat16)
        want_b = xb.gather(0, full_idx[:, :, None].expand(-1, -1, C))
        part[0][own] = want_b[own]
        assert torch.equal(rd.assemble_owned(part), want_b[None])
        # every slot is owned exactly once
        cnt = torch.zeros(t * N, dtype=torch.int64)
        cnt[slots] += 1
        dist.all_reduce(cnt)
        assert bool((cnt == 1).all())

        # ---- PivotKV split by KV head
        H, KVH, L, D = 8, 4, 48, 16
        q = torch.randn(1, H, L, D, generator=g)
        k = torch.randn(1, KVH, L, D, generator=g)
        _, _, b_full = op.pivot_scores(q, k, "cuda", return_partials=True)      # [KVH, L]
        per = [KVH // world] * world
        g0 = rank * per[0]
        G = H // KVH
        _, _, b_loc = op.pivot_scores(q[:, g0 * G:(g0 + per[0]) * G], k[:, g0:g0 + per[0]], "cuda", return_partials=True)
        hs = rd.gather_head_scores(b_loc, per)
        assert torch.equal(hs, b_full), "per-KV-head scores do not depend on which heads share a rank"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sync", [False, True])
def test_two_rank_host_logic(sync):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, sync), nprocs=2, join=True)


def _runner_worker(rank, world, port):
    for p in (ROOT, os.path.join(ROOT, "video-retake_b200")):
        sys.path.insert(0, p)
    from retake import infer_eval as ie
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        assert ie.shard_indices(7) == [i for i in range(7) if i % world == rank]
        seen = []
        merged = ie.run_sharded([10, 11, 12, 13, 14], lambda v: (seen.append(v), v * v)[1], ids="abcde")
        assert seen == [10 + i for i in range(5) if i % world == rank]            # whole videos, round robin
        assert merged == {"a": 100, "b": 121, "c": 144, "d": 169, "e": 196}       # same on every rank
        try:
            ie.gather_results({"dup": rank})
            raise AssertionError("a video processed twice must be reported")
        except RuntimeError:
            pass
    finally:
        dist.destroy_process_group()


def test_two_rank_whole_video_runner():
    """reference pattern infer_eval.py:181-205: round-robin shard, one all_gather_object of the results"""
    port = _free_port()
    mp.spawn(_runner_worker, args=(2, port), nprocs=2, join=True)


def test_frame_sampling_arithmetic():
    """demo.py:16-25 / dataset_utils.py:39-48 on hand-checked cases"""
    for p in (ROOT, os.path.join(ROOT, "video-retake_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from retake import infer_eval as ie
    # 1 h of video extracted at 25 fps, sampled at 2 fps, cap 2048 -> 2048 frames
    assert ie.get_sample_frames(90000, 2048, 2, 25) == 2048
    # 100 s at 25 fps sampled at 2 fps -> 200; odd counts round down to even; short clips are capped by their length
    assert ie.get_sample_frames(2500, 2048, 2, 25) == 200
    assert ie.get_sample_frames(2513, 2048, 2, 25) == 200 and ie.get_sample_frames(2538, 2048, 2, 25) == 202
    assert ie.get_sample_frames(7, 2048, 30, 25) == 6
    idx = ie.get_frame_indices(2500, 2048, 2, 25)
    assert idx.dtype.name == "int32" and len(idx) == 200 and idx[0] == 0 and idx[-1] == 2499
    assert idx.tolist()[:4] == [0, 12, 25, 37] and all(b > a for a, b in zip(idx, idx[1:]))
    # single process: run_sharded degenerates to a plain loop
    assert ie.run_sharded([1, 2, 3], lambda v: -v) == {0: -1, 1: -2, 2: -3}
    with pytest.raises(ValueError):
        ie.shard_indices(4, 2, 2)
