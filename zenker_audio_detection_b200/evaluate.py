"""Snippet-level scoring on the GPU: the batched entry point of the reference's evaluators.

The reference scores short labelled snippets in two places, both through the same two contracts as the long-audio path:

* ``utils/analyze_ROC_PR_stage1.py:163-191`` (and ``..._stage2.py``): ``run_inference(model_dir, X, batch_size)`` loads
  the extractor and the model from ``model_dir``, walks ``X`` in batches of ``batch_size`` (each entry a file path, a
  ``{"array", "sampling_rate"}`` dict or a waveform), and returns ``softmax(logits)[:, 1]`` as one float32 vector;
* ``src/test_trained_model_stage1_cv.py:127-162`` (and ``..._stage2_cv.py``): ``Trainer.predict`` over a dataset whose
  transform is the extractor; ``predictions.predictions`` are the logits, ``argmax(axis=1)`` the predicted classes.

Those scripts already run unmodified under ``compat.patch_transformers()`` (their per-batch Python loop then drives the
CUDA extractor and model).  The functions here remove the loop: snippets are decoded once, grouped by length (the
feature kernel takes equal-length rows), their features are computed per group straight into one ``(N, max_length,
128)`` device tensor in the caller's order, and the forward runs over large batches.  ``run_inference`` keeps the
reference's name, arguments and result; ``snippet_logits`` is what replaces ``Trainer.predict(...).predictions``.

No CPU fallback: everything below ends in ``ops`` calls, which raise without the CUDA library or an sm_100 device.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Any, List, Optional, Sequence

import numpy as np
import torch

SAMPLING_RATE = 16000


def to_waveform(entry: Any, device: Optional[torch.device] = None) -> torch.Tensor:
    """One snippet as a mono float32 16 kHz CUDA vector.  Accepts what ``to_waveform`` of the reference accepts
    (utils/analyze_ROC_PR_stage1.py:133-155): a waveform array, a dict payload, or a file path.  Resampling and the
    channel mean run on the GPU (``ops.resample`` = torchaudio's sinc kernel, ref:152-154)."""
    from . import ops, wavio

    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if isinstance(entry, torch.Tensor):
        entry = entry.detach().cpu().numpy()
    if isinstance(entry, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(entry, dtype=np.float32)).to(dev)
    if isinstance(entry, dict):
        arr = None
        for key in ("array", "audio", "values"):  # ref:137 (`or` chain; an ndarray has no truth value, so test for None)
            if entry.get(key) is not None:
                arr = entry[key]
                break
        if arr is None:
            raise ValueError("Unsupported dict payload for audio sample.")
        sr = entry.get("sampling_rate") or entry.get("sampling_rate_hz") or SAMPLING_RATE
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(arr, dtype=np.float32))).to(dev)
        return ops.resample(t, int(sr), SAMPLING_RATE) if int(sr) != SAMPLING_RATE else t
    if isinstance(entry, (str, bytes)) or hasattr(entry, "__fspath__"):
        samples, sr = wavio.read(entry)  # (channels, n) float32 or (n, channels) int16, as stored
        t = torch.from_numpy(np.ascontiguousarray(samples)).to(dev)
        return ops.resample(t, int(sr), SAMPLING_RATE).reshape(-1)  # channel mean fused (+ resample when needed)
    raise TypeError(f"Unsupported audio payload type: {type(entry)}")


def snippet_features(fx, entries: Sequence[Any], device: Optional[torch.device] = None) -> torch.Tensor:
    """``(N, max_length, num_mel_bins)`` float32 features of ``entries`` in the caller's order: what
    ``fx(wavs, sampling_rate=16000, return_tensors="pt")["input_values"]`` returns (HF:feature_extraction...:158-232),
    computed with one kernel launch per distinct snippet length instead of one per snippet."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    n = len(entries)
    out = torch.empty((n, fx.max_length, fx.num_mel_bins), dtype=torch.float32, device=dev)
    if n == 0:
        return out
    with torch.cuda.device(dev):
        plan = fx._get_plan(dev)
        groups = defaultdict(list)
        for i, e in enumerate(entries):
            w = to_waveform(e, dev)
            if w.ndim != 1:
                raise ValueError(f"Only mono-channel audio is supported for input to {fx}")
            groups[int(w.numel())].append((i, w))
        for _, members in groups.items():
            idx = torch.tensor([i for i, _ in members], dtype=torch.long, device=dev)
            feats = plan.fx_contract(torch.stack([w for _, w in members]), fx.mean, fx.std, fx.max_length, fx.do_normalize)
            out.index_copy_(0, idx, feats)
    return out


def snippet_logits(model, fx, entries: Sequence[Any], batch_size: int = 256) -> np.ndarray:
    """``(N, num_labels)`` float32 logits in the caller's order = ``Trainer.predict(dataset).predictions`` of
    src/test_trained_model_stage1_cv.py:156-159 (decisions re-checked like every other forward of the model)."""
    n = len(entries)
    if n == 0:
        return np.zeros((0, int(model.num_labels)), dtype=np.float32)
    dev = model.device if model.device.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
    model.to(dev)
    rows = []
    step = max(1, int(batch_size))
    for start in range(0, n, step):  # features of one forward batch at a time: 512 KiB per snippet
        feats = snippet_features(fx, entries[start:start + step], dev)
        rows.append(model(feats).logits)
    return torch.cat(rows, dim=0).float().cpu().numpy()


def run_inference(model_dir: str, X: List, batch_size: int) -> np.ndarray:
    """Same name, arguments and result as utils/analyze_ROC_PR_stage1.py:163-191: probability of class 1 per entry."""
    from .fx import ZenkerASTFeatureExtractor
    from .model import ZenkerASTForAudioClassification

    if len(X) == 0:
        return np.zeros((0,), dtype=np.float32)  # ref:191
    fx = ZenkerASTFeatureExtractor.from_pretrained(model_dir)
    model = ZenkerASTForAudioClassification.from_pretrained(model_dir).to("cuda").eval()
    logits = torch.from_numpy(snippet_logits(model, fx, X, batch_size))
    return torch.softmax(logits, dim=1)[:, 1].numpy()
