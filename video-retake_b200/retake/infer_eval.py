"""Whole-video data parallelism: the runner side of SURVEY.md 8(e) "whole videos" / 8(f4).

The reference evaluates a dataset with one process per GPU, a round-robin shard of the videos per rank and one
``all_gather_object`` of the per-video results at the end (``retake/infer_eval.py:159-211``); frames are sampled with
``get_frame_indices`` (``demo.py:16-25``, the same arithmetic as ``dataset_utils.py:39-48``).  This module keeps those
two pieces - the sampling arithmetic and the shard / gather loop - for a process group that already exists (``torchrun``
or ``bench.py --gpus N`` style launch, NCCL on the GPUs, gloo in the CPU tests); model loading, decoding of video files
and the benchmark-specific scoring stay out of scope.
"""
from __future__ import annotations

import math
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch.distributed as dist

__all__ = ["get_sample_frames", "get_frame_indices", "shard_indices", "gather_results", "run_sharded"]


def get_sample_frames(total_frames: int, max_num_frames: int, sample_fps: float, extraction_fps: float) -> int:
    """frames to sample from a video of ``total_frames`` extracted at ``extraction_fps`` when sampling at ``sample_fps``,
    capped by ``max_num_frames`` and rounded DOWN to an even count (temporal_patch_size 2): ``dataset_utils.py:39-48``"""
    sample_frames = float(total_frames / extraction_fps) * sample_fps
    sample_frames = min(total_frames, max_num_frames, sample_frames)
    sample_frames = math.floor(sample_frames)
    return int(sample_frames / 2) * 2


def get_frame_indices(total_frames: int, max_num_frames: int, sample_fps: float, extraction_fps: float) -> np.ndarray:
    """uniformly spaced frame indices (``np.linspace(0, total - 1, n).astype(int32)``): ``demo.py:16-25``"""
    n = get_sample_frames(total_frames, max_num_frames, sample_fps, extraction_fps)
    return np.linspace(0, total_frames - 1, n).astype(np.int32)


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_indices(n_items: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> List[int]:
    """round-robin shard of ``range(n_items)`` for ``rank`` (``infer_eval.py:181``)"""
    if rank is None or world_size is None:
        rank, world_size = _world()
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return [i for i in range(n_items) if i % world_size == rank]


def gather_results(local: Dict[Any, Any], group=None) -> Dict[Any, Any]:
    """merge the per-rank ``{item id: result}`` dicts on every rank (``infer_eval.py:196-205``); one object all-gather,
    no data-path collective"""
    rank, world = _world(group)
    if world == 1:
        return dict(local)
    parts: List[Optional[Dict[Any, Any]]] = [None] * world
    dist.barrier(group)
    dist.all_gather_object(parts, local, group=group)
    merged: Dict[Any, Any] = {}
    for d in parts:
        for k, v in d.items():
            if k in merged:
                raise RuntimeError(f"item {k!r} was processed by more than one rank")
            merged[k] = v
    return merged


def run_sharded(items: Sequence[Any], infer_fn: Callable[[Any], Any], ids: Optional[Iterable[Any]] = None,
                group=None) -> Dict[Any, Any]:
    """``infer_fn(item)`` on this rank's round-robin shard of ``items`` (one whole video per call - DPSelect, chunked
    prefill with PivotKV and decoding all stay on this rank's GPU), then the merged ``{id: result}`` on every rank"""
    rank, world = _world(group)
    ids = list(ids) if ids is not None else list(range(len(items)))
    if len(ids) != len(items):
        raise ValueError("ids and items must have the same length")
    local = {ids[i]: infer_fn(items[i]) for i in shard_indices(len(items), rank, world)}
    return gather_results(local, group)
