"""The reference's own numerics, executed on CPU (TEST INFRASTRUCTURE ONLY).

The reference scripts are glue around the *installed* ``torchaudio`` and
``transformers``; these packages are part of the image both here and on
the GPU box, so the oracle can call them directly.  Everything here is forced to CPU /
fp32 so the result does not depend on the box having a GPU.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import numpy as np
import torch


def resample(wave: np.ndarray, orig: int, new: int) -> np.ndarray:
    """``load_audio`` body after decoding: src/test_long_audio_windows_2stage.py:55-59."""
    import torchaudio

    w = torch.from_numpy(np.ascontiguousarray(wave, dtype=np.float32))
    if w.dim() == 1:
        w = w.unsqueeze(0)
    if w.size(0) > 1:
        w = w.mean(dim=0, keepdim=True)
    if orig != new:
        w = torchaudio.functional.resample(w, orig, new)
    return w.squeeze(0).numpy()


def kaldi_fbank(wave: np.ndarray, window_type: str = "hanning", num_mel_bins: int = 128, dtype=torch.float32) -> np.ndarray:
    import torchaudio.compliance.kaldi as K

    w = torch.from_numpy(np.ascontiguousarray(wave)).to(dtype).unsqueeze(0)
    return K.fbank(w, sample_frequency=16000, window_type=window_type, num_mel_bins=num_mel_bins).numpy()


def hf_feature_extractor(mean: float, std: float, max_length: int = 1024, num_mel_bins: int = 128):
    from transformers import ASTFeatureExtractor

    return ASTFeatureExtractor(mean=mean, std=std, max_length=max_length, num_mel_bins=num_mel_bins)


def hf_model_from_state_dict(sd, num_labels: int = 2):
    from transformers import ASTConfig, ASTForAudioClassification

    cfg = ASTConfig(num_labels=num_labels)
    with torch.device("cpu"):
        m = ASTForAudioClassification(cfg)
    m.load_state_dict(sd)
    return m.eval()


def forward_probs(model, fx, windows: Sequence[np.ndarray], batch_size: int, device="cpu") -> np.ndarray:
    """src/test_long_audio_windows_2stage.py:104-113 with DEVICE pinned to ``device``."""
    out: List[np.ndarray] = []
    with torch.inference_mode():
        for i in range(0, len(windows), batch_size):
            batch = list(windows[i : i + batch_size])
            inputs = fx(batch, sampling_rate=16000, return_tensors="pt")
            feats = inputs[fx.model_input_names[0]].to(device)
            logits = model(feats).logits
            out.append(torch.softmax(logits, dim=1).cpu().numpy())
    return np.concatenate(out, axis=0) if out else np.zeros((0,))


def set_threads(n: int | None = None) -> int:
    n = n or os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()
