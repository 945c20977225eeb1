"""Oracle restatement of the reference-owned glue (TEST INFRASTRUCTURE ONLY).

Every function cites the reference lines it follows; ``ref:`` means
``/root/reference/src/test_long_audio_windows_2stage.py`` and ``refc:`` the
``..._cache.py`` variant.  The restatement is vectorised numpy, the reference is
Python loops; results (values, dtypes, dict keys, None/NaN conventions) are the same
and are pinned by ``tests/golden/glue_*.json`` generated from the reference itself.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

SAMPLING_RATE = 16000  # ref:47


def window_starts(num_samples: int, window_sec: float, hop_sec: float, sr: int = SAMPLING_RATE) -> List[int]:
    """Start sample of every window.  ref:62-75 (``range(0, max(1, L-win+1), hop)``).

    No tail window is emitted; a recording shorter than one window yields exactly one
    (zero padded) window.
    """
    win = int(window_sec * sr)
    hop = int(hop_sec * sr)
    return list(range(0, max(1, num_samples - win + 1), hop))


def window_audio(audio: np.ndarray, window_sec: float, hop_sec: float, sr: int = SAMPLING_RATE) -> List[np.ndarray]:
    """ref:62-75.  Slices are views; only a too-short recording is zero padded."""
    win = int(window_sec * sr)
    out = []
    for s in window_starts(len(audio), window_sec, hop_sec, sr):
        seg = audio[s : s + win]
        if seg.shape[0] < win:
            seg = np.concatenate([seg, np.zeros(win - seg.shape[0], dtype=audio.dtype)])
        out.append(seg)
    return out


def stage1_gate(
    s1_probs: np.ndarray, stage1_threshold: float, forward_min_prob: Optional[float] = None
) -> Tuple[np.ndarray, np.ndarray]:
    """Stage-1 gate.  ref:312-320; refc:463-478 adds ``forward_min_prob``.

    Returns ``(s1_preds int64 (N,), swallow_indices int64 (K,))``.  ``argmax`` ties go
    to class 0 (numpy first-max rule), so ``pred==1`` iff ``p1 > p0``.
    """
    if s1_probs.ndim != 2 or s1_probs.shape[1] != 2:
        raise RuntimeError("Stage1 output shape unexpected; expected (N,2)")  # ref:310-311
    p_swallow = s1_probs[:, 1]
    arg = s1_probs.argmax(axis=1)
    preds = np.where((arg == 1) & (p_swallow >= stage1_threshold), 1, 0)
    idx = np.where(preds == 1)[0]
    if forward_min_prob is not None and len(idx):
        idx = idx[p_swallow[idx] >= forward_min_prob]
    return preds, idx


def stage2_classes(
    num_windows: int,
    stage2_results: Sequence[Tuple[int, np.ndarray]],
    stage2_threshold: float,
    use_argmax: bool = False,
) -> np.ndarray:
    """Per-window class vector: -1 idle, 0 healthy, 1 zenker.  ref:332-340; refc:510-522."""
    cls = np.full(num_windows, -1, dtype=int)
    for gidx, probs in stage2_results:
        if use_argmax:
            cls[gidx] = int(np.argmax(probs))
        else:
            cls[gidx] = 1 if probs[1] >= stage2_threshold else 0
    return cls


def summarize_stage_outputs(
    stage1_probs: np.ndarray,
    stage2_results: Sequence[Tuple[int, np.ndarray]],
    stage2_threshold: float = 0.5,
    use_argmax: bool = False,
) -> Dict[str, Any]:
    """ref:148-195 / refc:243-297.

    Reproduces the counting quirk: ``stage1_swallow_windows`` uses
    a BARE argmax, i.e. it ignores ``--stage1-threshold``; ``np.mean([])`` over an
    empty Stage-2 set yields NaN when the bare argmax count is non-zero.
    """
    n = len(stage1_probs)
    bare = stage1_probs.argmax(axis=1) if n else np.zeros((0,), dtype=np.int64)
    swallow_count = int((bare == 1).sum())
    idle_count = int((bare == 0).sum())
    aligned: List[Optional[np.ndarray]] = [None] * n
    for idx, probs in stage2_results:
        aligned[idx] = probs
    present = [p for p in aligned if p is not None]
    if use_argmax:
        zenker = sum(1 for p in present if np.argmax(p) == 1)
        healthy = sum(1 for p in present if np.argmax(p) == 0)
    else:
        zenker = sum(1 for p in present if p[1] >= stage2_threshold)
        healthy = sum(1 for p in present if p[1] < stage2_threshold)
    if swallow_count:
        with np.errstate(all="ignore"):
            import warnings

            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                mean2 = np.mean(present, axis=0)
        mean2 = mean2.tolist() if isinstance(mean2, np.ndarray) else float(mean2)
    else:
        mean2 = None
    return {
        "num_windows": int(n),
        "stage1_idle_windows": idle_count,
        "stage1_swallow_windows": swallow_count,
        "stage1_swallow_ratio": (swallow_count / n) if n else 0.0,
        "stage1_mean_probs": stage1_probs.mean(axis=0).tolist() if n else None,
        "stage2_mean_probs_over_swallow": mean2,
        "stage2_swallow_windows_evaluated": int(len(present)),
        "stage2_healthy_windows": int(healthy),
        "stage2_zenker_windows": int(zenker),
        "stage2_zenker_ratio_over_swallow": (zenker / swallow_count) if swallow_count else None,
    }


def aggregate_patient(per_file: Dict[str, Dict[str, Any]], files: List[str]) -> Dict[str, Any]:
    """ref:361-382."""
    vals = list(per_file.values())
    total_windows = int(sum(f["num_windows"] for f in vals))
    total_swallow = sum(f["stage1_swallow_windows"] for f in vals)
    total_zenker = sum(f["stage2_zenker_windows"] for f in vals)
    return {
        "files_used": files,
        "total_windows": total_windows,
        "total_idle_windows": int(sum(f["stage1_idle_windows"] for f in vals)),
        "total_swallow_windows": int(total_swallow),
        "total_swallow_ratio": total_swallow / max(1, total_windows),
        "total_swallow_windows_evaluated_stage2": int(sum(f["stage2_swallow_windows_evaluated"] for f in vals)),
        "total_healthy_windows": int(sum(f["stage2_healthy_windows"] for f in vals)),
        "total_zenker_windows": int(total_zenker),
        "overall_zenker_ratio_over_swallow": (total_zenker / total_swallow) if total_swallow else None,
    }


def cascade_file(
    s1_probs: np.ndarray,
    s2_probs_fn,
    stage1_threshold: float,
    stage2_threshold: float,
    forward_min_prob: Optional[float] = None,
    use_argmax: bool = False,
) -> Dict[str, Any]:
    """One file of ref:301-348 given Stage-1 probs and a callable ``idx -> (K,2)`` for Stage 2."""
    preds, idx = stage1_gate(s1_probs, stage1_threshold, forward_min_prob)
    results: List[Tuple[int, np.ndarray]] = []
    if len(idx):
        s2 = s2_probs_fn(idx)
        if s2.ndim != 2 or s2.shape[1] != 2:
            raise RuntimeError("Stage2 output shape unexpected; expected (K,2)")  # ref:325-326
        results = [(int(g), s2[i]) for i, g in enumerate(idx)]
    summary = summarize_stage_outputs(s1_probs, results, stage2_threshold, use_argmax)
    return {
        "s1_preds": preds,
        "swallow_indices": idx,
        "stage2_results": results,
        "classes": stage2_classes(len(preds), results, stage2_threshold, use_argmax),
        "summary": summary,
    }
