"""Multi-GPU layer of the two-stage path: recordings are independent (ref:301-348), so they are dealt to ranks with
no data-path collective; the only communication is the final gather of per-window score
records, one ``all_gather`` of counts followed by one ``all_gather`` of a padded record buffer (NCCL has no
allgatherv).  The same code runs on ``gloo`` for the CPU tests.

Record layout (24 bytes per window): six 32-bit words
    [recording_id i32, window_index i32, p_s1_0 f32, p_s1_1 f32, p_s2_0 f32, p_s2_1 f32]   (p_s2 = NaN if not forwarded)
carried as an int32 matrix (the float words bit-cast), so ids and probabilities both travel exactly.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

RECORD_WIDTH = 6  # 32-bit words


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


def shard_window_ranges(window_counts: Sequence[int], world_size: int,
                        weights: Optional[Sequence[float]] = None) -> List[List[Tuple[int, int, int]]]:
    """Finer than ``shard_recordings``: the windows of all recordings, laid end to end in recording order, are cut into
    ``world_size`` runs, so a rank gets ``(recording, first_window, end_window)`` chunks -- whole recordings plus at
    most one partial recording at either end of its run.  A window's scores do not depend on which other windows
    share its batch or its fbank (``TwoStagePipeline.run_audio16k(window_range=...)``; tested bit-exactly), so a split
    recording gives the same records as an unsplit one.  Runs are equal (to one window) by default; ``weights`` (one
    positive number per rank, e.g. the windows/s each GPU measured on a common probe -- the GPUs of a box settle at
    different power-capped clocks) makes them proportional instead.  Deterministic given its arguments: every rank
    must pass the same ``weights`` (all-gather them first)."""
    total = int(sum(int(c) for c in window_counts))
    if weights is None:
        bounds = [(r * total) // world_size for r in range(world_size + 1)]
    else:
        w = [float(v) for v in weights]
        if len(w) != world_size or not all(v > 0 and v == v and v != float("inf") for v in w):
            raise ValueError("weights: one positive finite number per rank")
        acc, cum = 0.0, [0.0]
        for v in w:
            acc += v
            cum.append(acc)
        bounds = [min(total, int(total * c / acc + 1e-9)) for c in cum]
        bounds[0], bounds[-1] = 0, total
        for r in range(1, world_size + 1):
            bounds[r] = max(bounds[r], bounds[r - 1])
    shards: List[List[Tuple[int, int, int]]] = [[] for _ in range(world_size)]
    base = 0
    for i, c in enumerate(window_counts):
        c = int(c)
        for r in range(world_size):
            lo, hi = max(bounds[r], base), min(bounds[r + 1], base + c)
            if lo < hi:
                shards[r].append((i, lo - base, hi - base))
        base += c
    return shards


def pack_records(recording_id: int, s1_probs: np.ndarray, swallow_indices: np.ndarray, s2_probs: np.ndarray,
                 window_base: int = 0) -> np.ndarray:
    """-> (n, 6) int32 record block of one recording, or of its windows ``window_base ...`` (a chunk of
    ``shard_window_ranges``; ``swallow_indices`` are relative to the chunk).  See the module docstring."""
    n = len(s1_probs)
    f = np.full((n, 4), np.nan, dtype=np.float32)
    f[:, 0:2] = s1_probs
    if len(swallow_indices):
        f[swallow_indices, 2:4] = s2_probs
    rec = np.empty((n, RECORD_WIDTH), dtype=np.int32)
    rec[:, 0] = recording_id
    rec[:, 1] = np.arange(window_base, window_base + n, dtype=np.int32)
    rec[:, 2:6] = f.view(np.int32)
    return rec


def unpack_records(rec: np.ndarray) -> Dict[int, Tuple[np.ndarray, np.ndarray, np.ndarray]]:
    """-> {recording_id: (s1_probs (N,2) f32, swallow_indices (K,) i64, s2_probs (K,2) f32)} in window order."""
    out: Dict[int, Tuple[np.ndarray, np.ndarray, np.ndarray]] = {}
    if len(rec) == 0:
        return out
    rec = np.ascontiguousarray(rec, dtype=np.int32)
    ids = rec[:, 0].astype(np.int64)
    for rid in np.unique(ids):
        r = rec[ids == rid]
        r = np.ascontiguousarray(r[np.argsort(r[:, 1], kind="stable")])
        f = r[:, 2:6].copy().view(np.float32)
        fwd = np.where(~np.isnan(f[:, 2]))[0].astype(np.int64)
        out[int(rid)] = (f[:, 0:2].copy(), fwd, f[fwd, 2:4].copy())
    return out


def all_gather_records(local: np.ndarray, device: torch.device) -> np.ndarray:
    """Gather every rank's (n_r, 6) int32 record block to every rank; returns them concatenated in rank order.
    Two collectives: the counts, then one padded buffer (NCCL has no allgatherv); 24 bytes per window on the wire."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    count = torch.tensor([local.shape[0]], dtype=torch.int64, device=device)
    counts = torch.zeros((world,), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(counts, count)
    counts_h = [int(c) for c in counts.cpu()]
    cap = max(max(counts_h), 1)
    buf = torch.zeros((cap, RECORD_WIDTH), dtype=torch.int32, device=device)
    if local.shape[0]:
        buf[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.int32)).to(device)
    gathered = torch.empty((world * cap, RECORD_WIDTH), dtype=torch.int32, device=device)  # rank blocks along dim 0
    dist.all_gather_into_tensor(gathered, buf)
    g = gathered.view(world, cap, RECORD_WIDTH).cpu().numpy()
    return np.concatenate([g[r, :c] for r, c in enumerate(counts_h)], axis=0)


def patient_documents(records: Dict[int, Tuple[np.ndarray, np.ndarray, np.ndarray]], patients: Sequence[Sequence[int]],
                      stage2_threshold: float, stage2_argmax: bool = False) -> List[Dict]:
    """ref:361-382 after the gather: per-file summaries (ref:148-195) and the aggregate of each patient's recordings."""
    from . import cascade

    docs = []
    for files in patients:
        names = [f"rec{int(i)}" for i in files]
        per_file = {}
        for name, i in zip(names, files):
            s1, idx, s2 = records[int(i)]
            per_file[name] = cascade.summarize_stage_outputs(s1, idx, s2, stage2_threshold, stage2_argmax)
        docs.append({"per_file": per_file, "aggregate": cascade.aggregate_patient(per_file, names)})
    return docs
