"""Constant tables of the audio front end, built on the host with the same torch fp32 ops the
reference's third-party numerics use, so they are bit-identical to torchaudio's:

* ``mel_banks``        TA:compliance/kaldi.py:436-511 (vtln_warp == 1 branch)
* ``feature_window``   TA:compliance/kaldi.py:86-111
* ``sinc_resample_kernel`` TA:functional/functional.py:1341-1402 (sinc_interp_hann, fp32 like the waveform)

These are tables, not the hot path; they are computed once per process and uploaded.
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import Tuple

import torch

EPSILON = torch.tensor(torch.finfo(torch.float).eps).item()  # TA:compliance/kaldi.py:22


@lru_cache(maxsize=None)
def mel_banks(num_bins: int = 128, padded: int = 512, sample_freq: float = 16000.0, low_freq: float = 20.0,
              high_freq: float = 0.0) -> torch.Tensor:
    """(num_bins, padded/2) fp32 triangular mel bank (no Nyquist column)."""
    num_fft_bins = padded / 2
    nyquist = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyquist
    fft_bin_width = sample_freq / padded
    mel_low = 1127.0 * math.log(1.0 + low_freq / 700.0)
    mel_high = 1127.0 * math.log(1.0 + high_freq / 700.0)
    delta = (mel_high - mel_low) / (num_bins + 1)
    b = torch.arange(num_bins).unsqueeze(1)
    left = mel_low + b * delta
    center = mel_low + (b + 1.0) * delta
    right = mel_low + (b + 2.0) * delta
    mel = (1127.0 * (1.0 + (fft_bin_width * torch.arange(num_fft_bins)) / 700.0).log()).unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    return torch.max(torch.zeros(1), torch.min(up, down)).contiguous()


@lru_cache(maxsize=None)
def feature_window(window_type: str = "hanning", size: int = 400) -> torch.Tensor:
    if window_type == "hanning":
        return torch.hann_window(size, periodic=False, dtype=torch.float32)
    if window_type == "hamming":
        return torch.hamming_window(size, periodic=False, alpha=0.54, beta=0.46, dtype=torch.float32)
    if window_type == "povey":
        return torch.hann_window(size, periodic=False, dtype=torch.float32).pow(0.85)
    if window_type == "rectangular":
        return torch.ones(size, dtype=torch.float32)
    raise ValueError(f"Invalid window type {window_type}")


@lru_cache(maxsize=None)
def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6,
                         rolloff: float = 0.99) -> Tuple[torch.Tensor, int, int, int]:
    """Returns ``(taps (new, 2*width+orig) fp32, width, orig, new)`` with orig/new reduced by their gcd."""
    if not (int(orig_freq) == orig_freq and int(new_freq) == new_freq):
        raise Exception("Frequencies must be of integer type to ensure quality resampling computation.")
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    dtype = torch.float32
    idx = torch.arange(-width, width + orig, dtype=dtype)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=dtype)[:, None, None] / new + idx
    t *= base_freq
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    scale = base_freq / orig
    kernels = torch.where(t == 0, torch.tensor(1.0).to(t), t.sin() / t)
    kernels *= window * scale
    return kernels.reshape(new, -1).contiguous(), width, orig, new
