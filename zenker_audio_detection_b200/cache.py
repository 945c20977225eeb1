"""Feature-cache compatibility with the reference's cached runner (refc = the reference's
``src/test_long_audio_windows_2stage_cache.py``).

The reference stores, per recording and per feature extractor, a ``torch.save`` bundle
``{"metadata": {...}, "features": (N, 1024, 128) float32 cpu}`` named ``<stem>_<sha256[:16]>.pt`` (refc:84-192).
This module reads and writes exactly that bundle, so caches written by either side are picked up by the other:

* key / file name / metadata checks follow refc:84-125,159-172 (same strings, same ``int(mtime)``),
* features that are not cached are computed by the CUDA extractor (``ZenkerASTFeatureExtractor``: fbank + pad +
  normalise kernels) instead of the CPU loop of refc:127-139,
* ``run_recording_cached`` is the per-recording flow of refc:433-507 on those bundles: Stage 1 from the cached
  features, gate + compaction on the GPU (``zk_gate_compact``), Stage 2 on ``index_select`` rows of the Stage-2
  features (the Stage-1 tensor again when both extractors are equal, refc:419-423).

The fused pipeline (``pipeline.TwoStagePipeline.run_waveform``) does not need a cache: it recomputes the continuous
fbank of a 10-minute recording in ~30 us.  This module exists for interoperability with cache directories that
reference runs have already filled, and so that a reference run can reuse features computed here.
"""
from __future__ import annotations

import hashlib
import json
import os
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import cascade, ops
from .pipeline import SAMPLING_RATE, RecordingResult, TwoStagePipeline


def get_fx_fingerprint(fx) -> str:
    """refc:84-86: sha256 of the sorted-key JSON of ``fx.to_dict()`` (ours equals HF's dict, tests/test_host_contract)."""
    return hashlib.sha256(json.dumps(fx.to_dict(), sort_keys=True).encode("utf-8")).hexdigest()


def build_cache_path(cache_dir: str, audio_path: str, window_sec: float, hop_sec: float, sr: int,
                     fx_fingerprint: str) -> str:
    """refc:89-104: ``<stem>_<first 16 hex of sha256(abs path|window|hop|sr|fingerprint|size_mtime)>.pt``."""
    audio_abs = os.path.abspath(audio_path)
    stats = f"{os.path.getsize(audio_abs)}_{int(os.path.getmtime(audio_abs))}"
    key = f"{audio_abs}|{window_sec}|{hop_sec}|{sr}|{fx_fingerprint}|{stats}"
    digest = hashlib.sha256(key.encode("utf-8")).hexdigest()[:16]
    stem = os.path.splitext(os.path.basename(audio_abs))[0]
    return os.path.join(cache_dir, f"{stem}_{digest}.pt")


def build_base_metadata(audio_path: str, window_sec: float, hop_sec: float, num_windows: int, sr: int,
                        fx_fingerprint: str) -> Dict[str, Any]:
    """refc:107-125: the keys a cached bundle must reproduce to be accepted."""
    audio_abs = os.path.abspath(audio_path)
    return {
        "audio_path": audio_abs,
        "audio_size": os.path.getsize(audio_abs),
        "audio_mtime": int(os.path.getmtime(audio_abs)),
        "window_sec": window_sec,
        "hop_sec": hop_sec,
        "num_windows": num_windows,
        "sampling_rate": sr,
        "extractor_fingerprint": fx_fingerprint,
    }


def compute_features(fx, windows: Union[Sequence[np.ndarray], torch.Tensor], batch_size: int) -> torch.Tensor:
    """refc:127-139 on the GPU: ``(N, max_length, num_mel_bins)`` float32 CUDA tensor for N one-second windows
    (a list of arrays as ``window_audio`` returns them, or an ``(N, samples)`` tensor)."""
    n = len(windows)
    if n == 0:
        raise RuntimeError("Feature extraction yielded no data; check window setup.")  # refc:135-136
    name = fx.model_input_names[0]
    chunks: List[torch.Tensor] = []
    for start in range(0, n, batch_size):
        batch = windows[start:start + batch_size]
        if torch.is_tensor(batch) and batch.is_cuda:  # windows already on the device: straight into the kernels
            chunks.append(fx._get_plan().fx_contract(batch.to(torch.float32).contiguous(), fx.mean, fx.std,
                                                     fx.max_length, fx.do_normalize))
            continue
        if not torch.is_tensor(batch):
            batch = list(batch)
        chunks.append(fx(batch, sampling_rate=SAMPLING_RATE, return_tensors="pt")[name])
    return torch.cat(chunks, dim=0).to(torch.float32).contiguous()


def load_bundle(cache_path: str, base_meta: Dict[str, Any]) -> Optional[torch.Tensor]:
    """refc:159-172: the cached features if the file exists, loads, and every key of ``base_meta`` matches."""
    if not os.path.exists(cache_path):
        return None
    bundle = torch.load(cache_path, map_location="cpu")
    metadata = bundle.get("metadata", {})
    if not all(metadata.get(k) == v for k, v in base_meta.items()):
        return None
    return bundle["features"].to(torch.float32).contiguous()


def save_bundle(cache_path: str, base_meta: Dict[str, Any], features: torch.Tensor) -> None:
    """refc:174-181: metadata + ``feature_shape``, features on the CPU."""
    meta = dict(base_meta)
    meta["feature_shape"] = list(features.shape)
    torch.save({"metadata": meta, "features": features.cpu()}, cache_path)


def load_or_compute_features(audio_path: str, windows, fx, window_sec: float, hop_sec: float, batch_size: int,
                             cache_dir: Optional[str], disable_cache: bool, refresh_cache: bool,
                             stage_label: str, log=print) -> torch.Tensor:
    """refc:142-192, same arguments and messages.  Returns a CPU tensor when the bundle was loaded and a CUDA tensor
    when the features were computed here (callers move batches with ``.to(device)`` either way, refc:204)."""
    fingerprint = get_fx_fingerprint(fx)
    base_meta = build_base_metadata(audio_path, window_sec, hop_sec, len(windows), SAMPLING_RATE, fingerprint)
    if disable_cache or not cache_dir:
        log(f"[cache:{stage_label}] Computing features (cache disabled).")
        return compute_features(fx, windows, batch_size)
    os.makedirs(cache_dir, exist_ok=True)
    cache_path = build_cache_path(cache_dir, audio_path, window_sec, hop_sec, SAMPLING_RATE, fingerprint)
    if not refresh_cache and os.path.exists(cache_path):
        try:
            features = load_bundle(cache_path, base_meta)
            if features is not None:
                log(f"[cache:{stage_label}] Loaded {cache_path}")
                return features
            log(f"[cache:{stage_label}] Metadata mismatch for {cache_path}; recomputing.")
        except Exception as exc:  # noqa: BLE001 - refc:171-172: best effort, fall through to recompute
            log(f"[cache:{stage_label}] Failed to load {cache_path}: {exc}; recomputing.")
    features = compute_features(fx, windows, batch_size)
    try:
        save_bundle(cache_path, base_meta, features)
        log(f"[cache:{stage_label}] Saved {cache_path}")
    except Exception as exc:  # noqa: BLE001 - refc:183-184
        log(f"[cache:{stage_label}] Failed to save {cache_path}: {exc}")
    return features


def window_audio(audio: Union[np.ndarray, torch.Tensor], window_sec: float, hop_sec: float) -> torch.Tensor:
    """ref:62-75 as one ``(N, win)`` strided view (no copy) of a 1-D 16 kHz recording; a recording shorter than one
    window gives a single zero-padded window."""
    a = torch.from_numpy(np.ascontiguousarray(audio)) if isinstance(audio, np.ndarray) else audio
    win, hop, n = cascade.window_geometry(int(a.numel()), window_sec, hop_sec)
    if a.numel() < win:
        a = torch.cat([a, a.new_zeros(win - a.numel())])
    return a.as_strided((n, win), (hop, 1))


def _logits_from_features(model, features: torch.Tensor, batch_size: int, device: torch.device,
                          rechecked: Optional[list] = None) -> torch.Tensor:
    """``model(batch).logits`` per batch (the drop-in ``__call__`` re-checks borderline windows itself, model.py)."""
    out = torch.empty((features.size(0), model.num_labels), dtype=torch.float32, device=device)
    with torch.inference_mode():
        for start in range(0, features.size(0), batch_size):
            batch = features[start:start + batch_size]
            if not batch.is_cuda:
                batch = (batch if batch.is_pinned() else batch.pin_memory()).to(device, non_blocking=True)
            out[start:start + batch.size(0)] = model(batch).logits
            if rechecked is not None:
                rechecked[0] += int(getattr(model, "last_rechecked", 0))
    return out


def forward_probs_from_features(model, features: torch.Tensor, batch_size: int) -> np.ndarray:
    """refc:198-208: ``(N, 2)`` float32 probabilities; an empty feature tensor gives ``np.zeros((0, 0))``."""
    if features.size(0) == 0:
        return np.zeros((0, 0))
    device = torch.device(model.device)
    with torch.cuda.device(device):
        return ops.softmax2(_logits_from_features(model, features, batch_size, device)).cpu().numpy()


def run_recording_cached(pipe: TwoStagePipeline, audio_path: str, windows, cache_dir: Optional[str],
                         disable_cache: bool = False, refresh_cache: bool = False, log=print) -> RecordingResult:
    """One recording through the cached flow of refc:433-531 with ``pipe``'s models, extractors and thresholds.

    ``windows``: what ``window_audio`` returns for the recording (list of ``(win,)`` float32 arrays) or the same as an
    ``(N, win)`` tensor.  The scores equal ``pipe.run_waveform`` up to the 16-bit forward's batch invariance; the gate,
    compaction, classes and summary are the same integer code paths.
    """
    n = len(windows)
    same_fx = pipe.fx1.__class__ is pipe.fx2.__class__ and pipe.fx1.to_dict() == pipe.fx2.to_dict()  # refc:419-421
    with torch.cuda.device(pipe.device):
        feats1 = load_or_compute_features(audio_path, windows, pipe.fx1, pipe.window_sec, pipe.hop_sec, pipe.batch_size,
                                          cache_dir, disable_cache, refresh_cache, "stage1", log)
        re1, re2 = [0], [0]
        logits1 = _logits_from_features(pipe.m1, feats1, pipe.batch_size, pipe.device, re1)
        if logits1.dim() != 2 or logits1.shape[1] != 2:
            raise RuntimeError("Stage1 output shape unexpected; expected (N,2)")  # refc:457-458
        probs1, pred, index, count = ops.gate_compact(logits1, pipe.thr1, pipe.min_prob)  # refc:459-475
        k = int(count.item())
        idx = index[:k].cpu().numpy().astype(np.int64)
        if k:
            feats2 = feats1 if same_fx else load_or_compute_features(
                audio_path, windows, pipe.fx2, pipe.window_sec, pipe.hop_sec, pipe.batch_size, cache_dir, disable_cache,
                refresh_cache, "stage2", log)
            picked = feats2.index_select(0, torch.as_tensor(idx, dtype=torch.long, device=feats2.device))  # refc:497-498
            logits2 = _logits_from_features(pipe.m2, picked, pipe.batch_size, pipe.device, re2)
            if logits2.shape[1] != 2:
                raise RuntimeError("Stage2 output shape unexpected; expected (K,2)")  # refc:502-503
            s2 = ops.softmax2(logits2).cpu().numpy()
        else:
            s2 = np.zeros((0, 2), dtype=np.float32)
        s1 = probs1.cpu().numpy()
        s1_preds = pred.cpu().numpy().astype(np.int64)
    classes = cascade.stage2_classes(n, idx, s2, pipe.thr2, pipe.stage2_argmax)
    summary = cascade.summarize_stage_outputs(s1, idx, s2, pipe.thr2, pipe.stage2_argmax)
    return RecordingResult(n, s1, s1_preds, idx, s2, classes, summary, re1[0], re2[0])
