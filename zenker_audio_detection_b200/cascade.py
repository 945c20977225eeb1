"""Host-side cascade bookkeeping of the two-stage path: window indexing, per-window classes, the per-file summary
and the per-patient aggregate.  Integer / dictionary logic only -- the numerics run on the GPU.

Follows ref: = /root/reference/src/test_long_audio_windows_2stage.py and refc: = ..._cache.py, including the
counting quirk of ``summarize_stage_outputs``: ``stage1_swallow_windows`` and the denominator of
``stage2_zenker_ratio_over_swallow`` use a BARE argmax that ignores ``--stage1-threshold``.
"""
from __future__ import annotations

import warnings
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

SAMPLING_RATE = 16000


def window_geometry(num_samples: int, window_sec: float, hop_sec: float, sr: int = SAMPLING_RATE) -> Tuple[int, int, int]:
    """(win, hop, n_windows) of ref:62-75: starts = range(0, max(1, L - win + 1), hop); no tail window; a
    recording shorter than one window yields exactly one zero-padded window."""
    win, hop = int(window_sec * sr), int(hop_sec * sr)
    if win <= 0 or hop <= 0:
        raise ValueError("window and hop must be positive")
    return win, hop, len(range(0, max(1, num_samples - win + 1), hop))


def window_starts(num_samples: int, window_sec: float, hop_sec: float, sr: int = SAMPLING_RATE) -> List[int]:
    win, hop, _ = window_geometry(num_samples, window_sec, hop_sec, sr)
    return list(range(0, max(1, num_samples - win + 1), hop))


def stage2_classes(num_windows: int, swallow_indices: np.ndarray, s2_probs: np.ndarray, stage2_threshold: float,
                   use_argmax: bool = False) -> np.ndarray:
    """-1 idle / 0 healthy / 1 zenker per window (ref:332-340; refc:510-522)."""
    cls = np.full(num_windows, -1, dtype=int)
    if len(swallow_indices):
        if use_argmax:
            z = s2_probs.argmax(axis=1) == 1
        else:
            z = s2_probs[:, 1] >= stage2_threshold
        cls[np.asarray(swallow_indices, dtype=np.int64)] = z.astype(int)
    return cls


def summarize_stage_outputs(stage1_probs: np.ndarray, swallow_indices: np.ndarray, s2_probs: np.ndarray,
                            stage2_threshold: float = 0.5, use_argmax: bool = False) -> Dict[str, Any]:
    """ref:148-195 / refc:243-297 for Stage-1 probabilities (N,2), the forwarded indices (K,) and their Stage-2
    probabilities (K,2)."""
    n = int(len(stage1_probs))
    bare = stage1_probs.argmax(axis=1) if n else np.zeros((0,), dtype=np.int64)
    swallow_count = int((bare == 1).sum())
    idle_count = int((bare == 0).sum())
    k = int(len(swallow_indices))
    if k:
        if use_argmax:
            am = s2_probs.argmax(axis=1)
            zenker, healthy = int((am == 1).sum()), int((am == 0).sum())
        else:
            zenker = int((s2_probs[:, 1] >= stage2_threshold).sum())
            healthy = int((s2_probs[:, 1] < stage2_threshold).sum())
    else:
        zenker = healthy = 0
    if swallow_count:
        if k:
            mean2: Any = np.mean([p for p in s2_probs], axis=0).tolist()
        else:  # ref:182-186: np.mean([], axis=0) -> nan (json.dump writes a bare NaN)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                mean2 = float(np.mean(np.zeros((0,)), axis=0))
    else:
        mean2 = None
    return {
        "num_windows": n,
        "stage1_idle_windows": idle_count,
        "stage1_swallow_windows": swallow_count,
        "stage1_swallow_ratio": (swallow_count / n) if n else 0.0,
        "stage1_mean_probs": stage1_probs.mean(axis=0).tolist() if n else None,
        "stage2_mean_probs_over_swallow": mean2,
        "stage2_swallow_windows_evaluated": k,
        "stage2_healthy_windows": healthy,
        "stage2_zenker_windows": zenker,
        "stage2_zenker_ratio_over_swallow": (zenker / swallow_count) if swallow_count else None,
    }


def aggregate_patient(per_file: Dict[str, Dict[str, Any]], files: List[str]) -> Dict[str, Any]:
    """ref:361-382."""
    vals = list(per_file.values())
    total_windows = int(sum(f["num_windows"] for f in vals))
    total_swallow = int(sum(f["stage1_swallow_windows"] for f in vals))
    total_zenker = int(sum(f["stage2_zenker_windows"] for f in vals))
    return {
        "files_used": list(files),
        "total_windows": total_windows,
        "total_idle_windows": int(sum(f["stage1_idle_windows"] for f in vals)),
        "total_swallow_windows": total_swallow,
        "total_swallow_ratio": total_swallow / max(1, total_windows),
        "total_swallow_windows_evaluated_stage2": int(sum(f["stage2_swallow_windows_evaluated"] for f in vals)),
        "total_healthy_windows": int(sum(f["stage2_healthy_windows"] for f in vals)),
        "total_zenker_windows": total_zenker,
        "overall_zenker_ratio_over_swallow": (total_zenker / total_swallow) if total_swallow else None,
    }


def classify_patient(aggregate: Dict[str, Any], threshold: float = 0.5) -> Optional[int]:
    """utils/aggregate_2stage_results.py:75-89: Zenker (1) iff overall ratio >= threshold, None when undefined."""
    r = aggregate.get("overall_zenker_ratio_over_swallow")
    if r is None:
        return None
    return 1 if r >= threshold else 0
