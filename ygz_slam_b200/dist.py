"""Multi-GPU plumbing for independent sequences: stream s -> rank s mod world, no data-path
collective; torch.distributed (NCCL on GPUs, gloo in the CPU tests) only gathers the per-stream result records and
reduces the counters / device times that the fps report needs."""
from __future__ import annotations

import numpy as np

RECORD_DOUBLES = 12  # stream id, frames, keypoints, matches, pose (7: t xyz + q wxyz), device ms


def shard_streams(n_streams: int, rank: int, world: int) -> list:
    """Streams owned by `rank`: round robin, every stream exactly once over all ranks."""
    return [s for s in range(n_streams) if s % world == rank]


def make_record(stream: int, frames: int, keypoints: int, matches: int, pose7, device_ms: float) -> np.ndarray:
    r = np.zeros(RECORD_DOUBLES, np.float64)
    r[0], r[1], r[2], r[3] = stream, frames, keypoints, matches
    r[4:11] = np.asarray(pose7, np.float64)
    r[11] = device_ms
    return r


def gather_records(local_records, n_streams: int, device=None):
    """all_gather of the per-stream records; returns (n_streams, RECORD_DOUBLES) sorted by stream id on every rank,
    plus (total frames, max device ms over ranks)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    per_rank = -(-n_streams // world)
    buf = torch.full((per_rank, RECORD_DOUBLES), -1.0, dtype=torch.float64, device=device)
    for i, r in enumerate(local_records):
        buf[i] = torch.from_numpy(np.asarray(r, np.float64)).to(buf.device)
    if world > 1:
        out = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(out, buf)
        allr = torch.cat(out).cpu().numpy()
    else:
        allr = buf.cpu().numpy()
    allr = allr[allr[:, 0] >= 0]
    allr = allr[np.argsort(allr[:, 0])]
    stats = torch.tensor([float(sum(r[1] for r in local_records)), float(sum(r[11] for r in local_records))],
                         dtype=torch.float64, device=device)
    frames, ms = stats.clone(), stats.clone()
    if world > 1:
        dist.all_reduce(frames, op=dist.ReduceOp.SUM)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return allr, float(frames[0]), float(ms[1])
