"""Seeded synthetic inputs and random-init AST weights.

No dataset or checkpoint is available offline, so every test and benchmark runs on
synthetic audio of the shapes BASELINE.json names and on random-init weights of the
AST-base architecture.  Generators are pure torch-CPU with explicit seeds, so the same
arrays are produced here and on the GPU box.

The random init is *conditioned* (SURVEY.md section 0.11): HF's ``_init_weights`` zeroes the
cls/distillation tokens and the position table (which hides token-order bugs) and gives a
window-to-window logit spread below bf16 noise; we randomise those tensors, give every
bias and LayerNorm a non-trivial value and scale the query/key projections by 4 so the
cascade gate has a real signal to split on.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch

HIDDEN, LAYERS, HEADS, MLP, PATCH, MEL, MAXLEN = 768, 12, 12, 3072, 16, 128, 1024
FSTRIDE = TSTRIDE = 10
NUM_TOKENS = ((MEL - PATCH) // FSTRIDE + 1) * ((MAXLEN - PATCH) // TSTRIDE + 1) + 2  # 1214
PFX = "audio_spectrogram_transformer."

# the reference's fallback stats (src/train_ast_stage1_cross_validation.py:104-105) for
# Stage 1 and a deliberately different pair for Stage 2 so the two extractors differ.
STAGE1_MEAN, STAGE1_STD = -1.1509622, 3.5340312
STAGE2_MEAN, STAGE2_STD = -2.0412, 4.1027


def random_state_dict(seed: int, num_labels: int = 2, qk_gain: float = 4.0, head_bias1: float = 0.0) -> Dict[str, torch.Tensor]:
    """All 203 tensors of an ``ASTForAudioClassification`` (HF key names, fp32, CPU)."""
    g = torch.Generator().manual_seed(seed)

    def n(*shape, std=0.02, mean=0.0):
        return torch.randn(*shape, generator=g) * std + mean

    sd: Dict[str, torch.Tensor] = {}
    e = PFX + "embeddings."
    sd[e + "cls_token"] = n(1, 1, HIDDEN)
    sd[e + "distillation_token"] = n(1, 1, HIDDEN)
    sd[e + "position_embeddings"] = n(1, NUM_TOKENS, HIDDEN)
    sd[e + "patch_embeddings.projection.weight"] = n(HIDDEN, 1, PATCH, PATCH)
    sd[e + "patch_embeddings.projection.bias"] = n(HIDDEN)
    for l in range(LAYERS):
        p = f"{PFX}encoder.layer.{l}."
        for nm, gain in (("query", qk_gain), ("key", qk_gain), ("value", 1.0)):
            sd[p + f"attention.attention.{nm}.weight"] = n(HIDDEN, HIDDEN) * gain
            sd[p + f"attention.attention.{nm}.bias"] = n(HIDDEN)
        sd[p + "attention.output.dense.weight"] = n(HIDDEN, HIDDEN)
        sd[p + "attention.output.dense.bias"] = n(HIDDEN)
        sd[p + "intermediate.dense.weight"] = n(MLP, HIDDEN)
        sd[p + "intermediate.dense.bias"] = n(MLP)
        sd[p + "output.dense.weight"] = n(HIDDEN, MLP)
        sd[p + "output.dense.bias"] = n(HIDDEN)
        for ln in ("layernorm_before", "layernorm_after"):
            sd[p + ln + ".weight"] = n(HIDDEN, std=0.1, mean=1.0)
            sd[p + ln + ".bias"] = n(HIDDEN, std=0.1)
    sd[PFX + "layernorm.weight"] = n(HIDDEN, std=0.1, mean=1.0)
    sd[PFX + "layernorm.bias"] = n(HIDDEN, std=0.1)
    sd["classifier.layernorm.weight"] = n(HIDDEN, std=0.1, mean=1.0)
    sd["classifier.layernorm.bias"] = n(HIDDEN, std=0.1)
    sd["classifier.dense.weight"] = n(num_labels, HIDDEN)
    b = torch.zeros(num_labels)
    b[1] = head_bias1
    sd["classifier.dense.bias"] = b
    return sd


def cfg1_windows(n: int = 64, seed: int = 1001) -> np.ndarray:
    """``n`` one-second 16 kHz windows whose level spans 2.5 decades; every third one carries
    a Gaussian-enveloped tone.  (n, 16000) float32 in [-1, 1]."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(16000, dtype=torch.float32) / 16000.0
    out = torch.empty(n, 16000)
    gains = 10.0 ** (torch.rand(n, generator=g) * 2.5 - 3.0)
    for i in range(n):
        w = gains[i] * torch.randn(16000, generator=g)
        if i % 3 == 0:
            w = w + 0.3 * torch.sin(2 * math.pi * (200.0 + 150.0 * i) * t) * torch.exp(-(((t - 0.5) / 0.1) ** 2))
        out[i] = w
    return out.clamp_(-1, 1).numpy()


def recording(seconds: float = 600.0, sr: int = 48000, seed: int = 2002, bursts_per_min: float = 9.0) -> np.ndarray:
    """One long mono recording: low-passed noise floor level-modulated per 5-s segment by
    10^U(-1,1), plus chirp bursts (0.4-0.8 s, 100->2000 Hz, Hann envelope).  float32 in [-1,1]."""
    g = torch.Generator().manual_seed(seed)
    n = int(round(seconds * sr))
    x = torch.randn(n + 4, generator=g) * 0.01
    x = (x[:-4] + x[1:-3] + x[2:-2] + x[3:-1] + x[4:]) / 5.0
    seg = 5 * sr
    nseg = (n + seg - 1) // seg
    lev = 10.0 ** (torch.rand(nseg, generator=g) * 2.0 - 1.0)
    x = x * lev.repeat_interleave(seg)[:n]
    nb = max(1, int(round(bursts_per_min * seconds / 60.0)))
    starts = torch.rand(nb, generator=g) * max(0.0, seconds - 1.0)
    durs = 0.4 + 0.4 * torch.rand(nb, generator=g)
    amps = 0.1 + 0.4 * torch.rand(nb, generator=g)
    for s, d, a in zip(starts.tolist(), durs.tolist(), amps.tolist()):
        i0, m = int(s * sr), int(d * sr)
        if i0 + m > n:
            m = n - i0
        if m <= 1:
            continue
        tt = torch.arange(m, dtype=torch.float32) / sr
        phase = 2 * math.pi * (100.0 * tt + 0.5 * (1900.0 / d) * tt * tt)
        x[i0 : i0 + m] += a * torch.sin(phase) * torch.hann_window(m, periodic=False)
    return x.clamp_(-1, 1).numpy()


def noise_16k(seconds: float = 3600.0, seed: int = 3003, level: float = 0.05) -> np.ndarray:
    """cfg3: plain Gaussian noise at 16 kHz for the fbank-throughput case."""
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(int(seconds * 16000), generator=g) * level).clamp_(-1, 1).numpy()
