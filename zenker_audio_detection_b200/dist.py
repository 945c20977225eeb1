"""Multi-GPU layer of the two-stage path: recordings are independent (ref:301-348), so they are dealt to ranks with
no data-path collective; the only communication is the final gather of per-window score
records, one ``all_gather`` of counts followed by one ``all_gather`` of a padded record buffer (NCCL has no
allgatherv).  The same code runs on ``gloo`` for the CPU tests.

Record layout (float64 x 7 per window, exact for int32 ids and float32 probabilities):
    [recording_id, window_index, p_s1_0, p_s1_1, forwarded (0/1), p_s2_0, p_s2_1]  (p_s2 = NaN if not forwarded)
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

RECORD_WIDTH = 7


def shard_recordings(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first: sort by length descending, always give the next recording to the least loaded
    rank (ties -> lowest rank).  Deterministic, independent of the calling rank."""
    loads = [0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += int(lengths[i])
    for s in shards:
        s.sort()
    return shards


def pack_records(recording_id: int, s1_probs: np.ndarray, swallow_indices: np.ndarray, s2_probs: np.ndarray) -> np.ndarray:
    n = len(s1_probs)
    rec = np.full((n, RECORD_WIDTH), np.nan, dtype=np.float64)
    rec[:, 0] = recording_id
    rec[:, 1] = np.arange(n)
    rec[:, 2:4] = s1_probs
    rec[:, 4] = 0.0
    if len(swallow_indices):
        rec[swallow_indices, 4] = 1.0
        rec[swallow_indices, 5:7] = s2_probs
    return rec


def unpack_records(rec: np.ndarray) -> Dict[int, Tuple[np.ndarray, np.ndarray, np.ndarray]]:
    """-> {recording_id: (s1_probs (N,2) f32, swallow_indices (K,) i64, s2_probs (K,2) f32)} in window order."""
    out: Dict[int, Tuple[np.ndarray, np.ndarray, np.ndarray]] = {}
    if len(rec) == 0:
        return out
    ids = rec[:, 0].astype(np.int64)
    for rid in np.unique(ids):
        r = rec[ids == rid]
        r = r[np.argsort(r[:, 1], kind="stable")]
        fwd = np.where(r[:, 4] == 1.0)[0].astype(np.int64)
        out[int(rid)] = (r[:, 2:4].astype(np.float32), fwd, r[fwd, 5:7].astype(np.float32))
    return out


def all_gather_records(local: np.ndarray, device: torch.device) -> np.ndarray:
    """Gather every rank's (n_r, 7) record block to every rank; returns them concatenated in rank order."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    count = torch.tensor([local.shape[0]], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count)
    counts_h = [int(c.item()) for c in counts]
    cap = max(max(counts_h), 1)
    buf = torch.zeros((cap, RECORD_WIDTH), dtype=torch.float64, device=device)
    if local.shape[0]:
        buf[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(device)
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)
    return np.concatenate([g[:c].cpu().numpy() for g, c in zip(gathered, counts_h)], axis=0)
