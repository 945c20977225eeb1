"""Oracle restatement of the third-party numerics on the path (TEST INFRASTRUCTURE ONLY).

The arithmetic of the reference's hot path lives in packages that are NOT under
``/root/reference``: torchaudio 2.11.0 (``functional.resample``,
``compliance.kaldi.fbank``) and transformers 5.5.0 (``ASTFeatureExtractor``,
``ASTForAudioClassification``); the reference pins only lower bounds
(``requirements.txt:2-12``).  This file restates their published algorithms:

* ``TA:`` = torchaudio/  ``HF:`` = transformers/models/audio_spectrogram_transformer/

numpy for the audio front end (dtype selectable: float32 mirrors the reference,
float64 is the "truth" used to bound both implementations' error), torch fp32 for the
transformer (this is the "plain fp32 reference" of the floating-point kernels).
``tests/test_oracle_numerics.py`` pins every function here against the installed
packages and against ``tests/golden/``.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

# ----------------------------------------------------------------------------------
# resample  (TA:functional/functional.py:1305-1432)
# ----------------------------------------------------------------------------------


def sinc_resample_kernel(
    orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99, dtype=np.float32
) -> Tuple[np.ndarray, int, int, int]:
    """Hann-windowed sinc polyphase taps.  TA:functional/functional.py:1341-1402.

    Returns ``(taps (new, 2*width+orig), width, orig, new)`` with orig/new reduced by gcd.
    """
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = np.arange(-width, width + orig, dtype=dtype)[None, :] / dtype(orig)
    t = np.arange(0, -new, -1, dtype=dtype)[:, None] / dtype(new) + idx
    t = t * dtype(base)
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width).astype(dtype)
    window = np.cos(t * dtype(math.pi) / dtype(lowpass_filter_width) / dtype(2)) ** 2
    t = t * dtype(math.pi)
    scale = dtype(base / orig)
    with np.errstate(divide="ignore", invalid="ignore"):
        taps = np.where(t == 0, dtype(1.0), np.sin(t) / t)
    taps = (taps * window * scale).astype(dtype)
    return taps, width, orig, new


def resample(wave: np.ndarray, orig_freq: int, new_freq: int, dtype=np.float32) -> np.ndarray:
    """1-D resample.  TA:functional/functional.py:1405-1432 (pad, strided conv, truncate)."""
    if orig_freq == new_freq:
        return wave.astype(dtype)
    taps, width, orig, new = sinc_resample_kernel(orig_freq, new_freq, dtype=dtype)
    n = wave.shape[-1]
    padded = np.concatenate([np.zeros(width, dtype), wave.astype(dtype), np.zeros(width + orig, dtype)])
    klen = taps.shape[1]
    n_steps = (padded.shape[0] - klen) // orig + 1
    frames = np.lib.stride_tricks.as_strided(
        padded, shape=(n_steps, klen), strides=(padded.strides[0] * orig, padded.strides[0])
    )
    out = (frames @ taps.T).reshape(-1)  # (n_steps, new) -> interleaved
    target = int(math.ceil(new * n / orig))
    return out[:target].astype(dtype)


# ----------------------------------------------------------------------------------
# kaldi fbank  (TA:compliance/kaldi.py)
# ----------------------------------------------------------------------------------

EPS_F32 = float(np.finfo(np.float32).eps)  # TA:compliance/kaldi.py:22 (1.1920929e-07)


def mel_scale(f):
    return 1127.0 * np.log(1.0 + f / 700.0)  # TA:compliance/kaldi.py:326-331


def mel_banks(num_bins: int = 128, padded: int = 512, sr: float = 16000.0, low: float = 20.0, high: float = 0.0,
              dtype=np.float64) -> np.ndarray:
    """HTK-style triangular mel bank, (num_bins, padded/2).  TA:compliance/kaldi.py:436-511
    (vtln_warp == 1 branch)."""
    nyq = 0.5 * sr
    if high <= 0:
        high += nyq
    nfft = padded // 2
    width = sr / padded
    mlo, mhi = 1127.0 * math.log(1.0 + low / 700.0), 1127.0 * math.log(1.0 + high / 700.0)
    delta = (mhi - mlo) / (num_bins + 1)
    b = np.arange(num_bins, dtype=dtype)[:, None]
    left, center, right = mlo + b * delta, mlo + (b + 1.0) * delta, mlo + (b + 2.0) * delta
    mel = mel_scale(width * np.arange(nfft, dtype=dtype))[None, :].astype(dtype)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    return np.maximum(0.0, np.minimum(up, down)).astype(dtype)


def feature_window(window_type: str, size: int = 400, dtype=np.float64) -> np.ndarray:
    """TA:compliance/kaldi.py:86-111 (hanning / hamming / povey / rectangular)."""
    n = np.arange(size, dtype=np.float64)
    hann = 0.5 - 0.5 * np.cos(2.0 * math.pi * n / (size - 1))
    if window_type == "hanning":
        w = hann
    elif window_type == "povey":
        w = hann ** 0.85
    elif window_type == "hamming":
        w = 0.54 - 0.46 * np.cos(2.0 * math.pi * n / (size - 1))
    elif window_type == "rectangular":
        w = np.ones(size)
    else:
        raise ValueError(window_type)
    return w.astype(dtype)


def num_frames(n: int, size: int = 400, shift: int = 160) -> int:
    return 0 if n < size else 1 + (n - size) // shift  # TA:compliance/kaldi.py:63-67


def fbank(
    wave: np.ndarray,
    window_type: str = "hanning",
    num_mel_bins: int = 128,
    sr: int = 16000,
    preemph: float = 0.97,
    dtype=np.float64,
) -> np.ndarray:
    """``kaldi.fbank`` with the defaults HF uses.  TA:compliance/kaldi.py:514-645 with
    ``dither=0, remove_dc_offset=True, snip_edges=True, use_power=True, use_log_fbank=True,
    use_energy=False, low_freq=20, high_freq=0`` (:516-540); HF call site
    HF:feature_extraction_audio_spectrogram_transformer.py:116-121 (hanning, no 2^15 scale).
    """
    size, shift = int(sr * 0.025), int(sr * 0.010)
    padded = 1 << (size - 1).bit_length()
    m = num_frames(wave.shape[0], size, shift)
    if m == 0:
        return np.zeros((0, num_mel_bins), dtype)
    x = wave.astype(dtype)
    fr = np.lib.stride_tricks.as_strided(x, shape=(m, size), strides=(x.strides[0] * shift, x.strides[0])).copy()
    fr -= fr.mean(axis=1, keepdims=True)  # :183-186
    prev = np.concatenate([fr[:, :1], fr[:, :-1]], axis=1)  # replicate pad :193-198
    fr = fr - dtype(preemph) * prev
    fr = fr * feature_window(window_type, size, dtype)[None, :]  # :200-204
    fr = np.concatenate([fr, np.zeros((m, padded - size), dtype)], axis=1)  # :207-211
    spec = np.abs(np.fft.rfft(fr.astype(np.float64), axis=1)).astype(dtype) ** 2  # :616-618
    bank = np.concatenate([mel_banks(num_mel_bins, padded, float(sr), dtype=dtype),
                           np.zeros((num_mel_bins, 1), dtype)], axis=1)  # :621-627
    mel = spec @ bank.T  # :630
    return np.log(np.maximum(mel, dtype(EPS_F32))).astype(dtype)  # :631-633


def fx_features(
    waves: Sequence[np.ndarray],
    mean: float,
    std: float,
    max_length: int = 1024,
    num_mel_bins: int = 128,
    do_normalize: bool = True,
    dtype=np.float32,
) -> np.ndarray:
    """``ASTFeatureExtractor.__call__`` body.  HF:feature_extraction...:104-156,215-227:
    per-waveform fbank, zero-pad / truncate to ``max_length`` rows, THEN normalise
    ``(x - mean) / (2*std)`` (so pad rows hold ``-mean/(2*std)``)."""
    out = np.zeros((len(waves), max_length, num_mel_bins), dtype)
    for i, w in enumerate(waves):
        fb = fbank(np.asarray(w, dtype=np.float32), num_mel_bins=num_mel_bins, dtype=dtype)
        k = min(max_length, fb.shape[0])
        out[i, :k] = fb[:k]
    if do_normalize:
        out = (out - dtype(mean)) / dtype(std * 2)
    return out.astype(dtype)


# ----------------------------------------------------------------------------------
# AST forward, torch fp32  (HF:modeling_audio_spectrogram_transformer.py)
# ----------------------------------------------------------------------------------

PFX = "audio_spectrogram_transformer."


def ast_forward(
    sd: Dict[str, torch.Tensor],
    input_values: torch.Tensor,
    num_layers: int = 12,
    num_heads: int = 12,
    eps: float = 1e-12,
    fstride: int = 10,
    tstride: int = 10,
    return_hidden: bool = False,
):
    """Logits of ``ASTForAudioClassification``.  HF:modeling...:62-72 (embeddings), :92-96
    (patch conv, frequency-major flatten), :150-181 (attention, scale d^-1/2, no mask),
    :263-281 (pre-LN block), :376-382 (final LN, mean of tokens 0 and 1), :385-394 (head).
    ``sd`` is the HF state dict (203 tensors); everything runs in fp32 on CPU."""
    F = torch.nn.functional
    x = input_values.to(torch.float32)
    B = x.shape[0]
    w = sd[PFX + "embeddings.patch_embeddings.projection.weight"].float()
    b = sd[PFX + "embeddings.patch_embeddings.projection.bias"].float()
    x = F.conv2d(x.unsqueeze(1).transpose(2, 3), w, b, stride=(fstride, tstride))  # (B,768,12,101)
    x = x.flatten(2).transpose(1, 2)
    D = x.shape[-1]
    cls = sd[PFX + "embeddings.cls_token"].float().expand(B, -1, -1)
    dist = sd[PFX + "embeddings.distillation_token"].float().expand(B, -1, -1)
    x = torch.cat([cls, dist, x], dim=1) + sd[PFX + "embeddings.position_embeddings"].float()
    dh = D // num_heads
    for l in range(num_layers):
        p = f"{PFX}encoder.layer.{l}."
        h = F.layer_norm(x, (D,), sd[p + "layernorm_before.weight"].float(), sd[p + "layernorm_before.bias"].float(), eps)
        q = F.linear(h, sd[p + "attention.attention.query.weight"].float(), sd[p + "attention.attention.query.bias"].float())
        k = F.linear(h, sd[p + "attention.attention.key.weight"].float(), sd[p + "attention.attention.key.bias"].float())
        v = F.linear(h, sd[p + "attention.attention.value.weight"].float(), sd[p + "attention.attention.value.bias"].float())
        q, k, v = (t.view(B, -1, num_heads, dh).transpose(1, 2) for t in (q, k, v))
        a = torch.softmax((q @ k.transpose(2, 3)) * (dh ** -0.5), dim=-1) @ v
        a = a.transpose(1, 2).reshape(B, -1, D)
        x = x + F.linear(a, sd[p + "attention.output.dense.weight"].float(), sd[p + "attention.output.dense.bias"].float())
        h = F.layer_norm(x, (D,), sd[p + "layernorm_after.weight"].float(), sd[p + "layernorm_after.bias"].float(), eps)
        h = F.gelu(F.linear(h, sd[p + "intermediate.dense.weight"].float(), sd[p + "intermediate.dense.bias"].float()))
        x = x + F.linear(h, sd[p + "output.dense.weight"].float(), sd[p + "output.dense.bias"].float())
    hidden = x
    x = F.layer_norm(x, (D,), sd[PFX + "layernorm.weight"].float(), sd[PFX + "layernorm.bias"].float(), eps)
    pooled = (x[:, 0] + x[:, 1]) / 2
    pooled = F.layer_norm(pooled, (D,), sd["classifier.layernorm.weight"].float(), sd["classifier.layernorm.bias"].float(), eps)
    logits = F.linear(pooled, sd["classifier.dense.weight"].float(), sd["classifier.dense.bias"].float())
    return (logits, hidden) if return_hidden else logits
