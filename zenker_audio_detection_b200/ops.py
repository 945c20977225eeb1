"""Tensor-level wrappers of the C ABI: allocate outputs with torch, pass ``data_ptr()`` and the current stream.

Every function requires CUDA tensors on an sm_100 device and raises ``ZkError`` otherwise; nothing here
computes on the CPU or through PyTorch operators.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Tuple, Union

import torch

from . import _lib, tables
from ._lib import ZkError, check

HID, MLP, HEADS = 768, 3072, 12


def _cuda(t: torch.Tensor, dtype: torch.dtype, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ZkError(f"{name}: expected a CUDA tensor (zenker-b200 has no CPU path)")
    if t.dtype != dtype:
        raise ZkError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------------------------------ resample
_taps_cache: Dict[Tuple[int, int, int], torch.Tensor] = {}


def _device_taps(orig_freq: int, new_freq: int, device: torch.device):
    taps, width, orig, new = tables.sinc_resample_kernel(int(orig_freq), int(new_freq))
    key = (orig, new, device.index if device.index is not None else torch.cuda.current_device())
    if key not in _taps_cache:
        _taps_cache[key] = taps.to(device)
    return _taps_cache[key], width, orig, new


def resample(waveform: torch.Tensor, orig_freq: int, new_freq: int) -> torch.Tensor:
    """``load_audio`` after decoding (ref:55-58): channel mean + ``torchaudio.functional.resample``.

    waveform: CUDA float32 ``(channels, n)`` or ``(n,)``; or CUDA int16 ``(n, channels)`` interleaved PCM.
    Returns CUDA float32 ``(ceil(n*new/orig),)``.
    """
    lib = _lib.load()
    _lib.require_device()
    if not waveform.is_cuda:
        raise ZkError("resample: expected a CUDA tensor")
    taps, width, orig, new = _device_taps(orig_freq, new_freq, waveform.device)
    if waveform.dtype == torch.int16:
        w = waveform.contiguous()
        if w.dim() == 1:
            w = w.unsqueeze(1)
        n, ch = w.shape
        if orig == new:  # same rate: only the 2^-15 scaling and the channel mean remain (identity tap)
            taps, width = torch.ones(1, 1, dtype=torch.float32, device=w.device), 0
        n_out = (n * new + orig - 1) // orig
        out = torch.empty(n_out, dtype=torch.float32, device=w.device)
        check(lib.zk_resample_pcm16(w.data_ptr(), n, ch, taps.data_ptr(), orig, new, width, out.data_ptr(), n_out,
                                    _lib.stream_ptr()), "zk_resample_pcm16")
        return out
    w = _cuda(waveform, torch.float32, "resample")
    if w.dim() == 1:
        w = w.unsqueeze(0)
    ch, n = w.shape
    if orig == new:  # same rate: the reference skips resample (ref:57); only the channel mean remains
        taps1 = torch.ones(1, 1, dtype=torch.float32, device=w.device)
        out = torch.empty(n, dtype=torch.float32, device=w.device)
        check(lib.zk_resample_f32(w.data_ptr(), n, ch, n, taps1.data_ptr(), 1, 1, 0, out.data_ptr(), n,
                                  _lib.stream_ptr()), "zk_resample_f32")
        return out
    n_out = (n * new + orig - 1) // orig
    out = torch.empty(n_out, dtype=torch.float32, device=w.device)
    check(lib.zk_resample_f32(w.data_ptr(), n, ch, n, taps.data_ptr(), orig, new, width, out.data_ptr(), n_out,
                              _lib.stream_ptr()), "zk_resample_f32")
    return out


# ------------------------------------------------------------------------------------------ fbank
class FbankPlan:
    """Device tables for the Kaldi fbank kernel (window, FFT twiddles, sparse mel bank)."""

    def __init__(self, window_type: str = "hanning", num_mel_bins: int = 128, preemph: float = 0.97,
                 sample_frequency: int = 16000):
        lib = _lib.load()
        _lib.require_device()
        if sample_frequency != 16000:
            raise ZkError("FbankPlan: only 16 kHz (25 ms = 400 / 10 ms = 160 samples) is supported")
        self.window_type, self.num_mel_bins, self.preemph = window_type, num_mel_bins, preemph
        win = tables.feature_window(window_type, 400).contiguous()
        mel = tables.mel_banks(num_mel_bins, 512, 16000.0, 20.0, 0.0).contiguous()
        h = C.c_void_p()
        check(lib.zk_fbank_plan_create(win.data_ptr(), mel.data_ptr(), num_mel_bins, preemph, tables.EPSILON,
                                       C.byref(h)), "zk_fbank_plan_create")
        self._h = h
        self._lib = lib

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.zk_fbank_plan_destroy(h)
            self._h = None

    @staticmethod
    def num_frames(n: int) -> int:
        return 0 if n < 400 else 1 + (n - 400) // 160

    def fbank(self, wave: torch.Tensor) -> torch.Tensor:
        """Continuous fbank: CUDA float32 ``(n,)`` -> ``(m, 128)``."""
        w = _cuda(wave, torch.float32, "fbank").reshape(-1)
        m = self.num_frames(w.numel())
        out = torch.empty((m, 128), dtype=torch.float32, device=w.device)
        if m:
            check(self._lib.zk_fbank_f32(self._h, w.data_ptr(), w.numel(), out.data_ptr(), m, _lib.stream_ptr()),
                  "zk_fbank_f32")
        return out

    def fx_contract(self, windows: torch.Tensor, mean: float, std: float, max_length: int,
                    do_normalize: bool = True) -> torch.Tensor:
        """``ASTFeatureExtractor.__call__`` body: CUDA float32 ``(B, win_len)`` -> ``(B, max_length, 128)``."""
        w = _cuda(windows, torch.float32, "fx_contract")
        if w.dim() != 2:
            raise ZkError("fx_contract: expected (batch, samples)")
        b, n = w.shape
        out = torch.empty((b, max_length, 128), dtype=torch.float32, device=w.device)
        if b:
            check(self._lib.zk_fx_contract_f32(self._h, w.data_ptr(), b, n, n, 1 if do_normalize else 0, float(mean),
                                               float(std), max_length, out.data_ptr(), _lib.stream_ptr()),
                  "zk_fx_contract_f32")
        return out


# ------------------------------------------------------------------------------------------ building blocks
def gemm(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, epilogue: int, out: Optional[torch.Tensor] = None,
         aux: Optional[torch.Tensor] = None, aux_rows: int = 0) -> torch.Tensor:
    lib = _lib.load()
    a = _cuda(a, torch.bfloat16, "gemm a")
    w = _cuda(w, torch.bfloat16, "gemm w")
    bias = _cuda(bias, torch.float32, "gemm bias")
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        if epilogue in (_lib.EPI_BIAS_BF16, _lib.EPI_BIAS_GELU_BF16):
            out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
        else:
            raise ZkError("gemm: the fp32 epilogues accumulate into / scatter to a caller-provided `out`")
    check(lib.zk_gemm_bf16(a.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, epilogue,
                           aux.data_ptr() if aux is not None else None, aux_rows, _lib.stream_ptr()), "zk_gemm_bf16")
    return out


def layernorm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float) -> torch.Tensor:
    lib = _lib.load()
    x = _cuda(x, torch.float32, "layernorm x")
    rows, cols = x.shape
    out = torch.empty((rows, cols), dtype=torch.bfloat16, device=x.device)
    check(lib.zk_layernorm_bf16(x.data_ptr(), _cuda(w, torch.float32, "w").data_ptr(),
                                _cuda(b, torch.float32, "b").data_ptr(), eps, out.data_ptr(), rows, cols,
                                _lib.stream_ptr()), "zk_layernorm_bf16")
    return out


def attention(qkv: torch.Tensor, batch: int, tokens: int) -> torch.Tensor:
    lib = _lib.load()
    qkv = _cuda(qkv, torch.bfloat16, "attention qkv")
    if qkv.shape != (batch * tokens, 3 * HID):
        raise ZkError(f"attention: qkv must be ({batch * tokens}, {3 * HID}), got {tuple(qkv.shape)}")
    out = torch.empty((batch * tokens, HID), dtype=torch.bfloat16, device=qkv.device)
    check(lib.zk_attention_bf16(qkv.data_ptr(), out.data_ptr(), batch, tokens, _lib.stream_ptr()), "zk_attention_bf16")
    return out


class FeatureStats:
    """Running mean / std of un-normalised features, as utils/compute_ast_normalization_stats.py:55-95 computes them
    (float64 sums over every element of the zero-padded ``(B, max_length, 128)`` tensors; unbiased std)."""

    def __init__(self, device: Optional[Union[str, torch.device]] = None):
        self.acc = torch.zeros(2, dtype=torch.float64, device=device or torch.device("cuda", torch.cuda.current_device()))
        self.count = 0

    def update(self, feats: torch.Tensor, padded_elements: Optional[int] = None) -> None:
        """``feats``: CUDA float32, any shape.  ``padded_elements``: element count of the padded tensor these values
        stand for (zero padding adds nothing to the sums; e.g. a compact (98,128) window counts as 1024*128)."""
        lib = _lib.load()
        f = _cuda(feats, torch.float32, "FeatureStats.update").reshape(-1)
        check(lib.zk_sum_sumsq_f64(f.data_ptr(), f.numel(), self.acc.data_ptr(), _lib.stream_ptr()), "zk_sum_sumsq_f64")
        self.count += int(padded_elements if padded_elements is not None else f.numel())

    def result(self) -> dict:
        if self.count == 0:
            return {"mean": 0.0, "std": 0.0, "count": 0}
        s, q = (float(v) for v in self.acc.cpu())
        mean = s / self.count
        var = max(q / self.count - mean * mean, 0.0)
        var = var * (self.count / (self.count - 1)) if self.count > 1 else 0.0
        return {"mean": float(mean), "std": float(var ** 0.5), "count": self.count}


def softmax2(logits: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    logits = _cuda(logits, torch.float32, "softmax2")
    n = logits.shape[0]
    probs = torch.empty_like(logits)
    if n:
        check(lib.zk_softmax2(logits.data_ptr(), n, probs.data_ptr(), _lib.stream_ptr()), "zk_softmax2")
    return probs


def gate_compact(logits: torch.Tensor, threshold: float, min_prob: Optional[float] = None):
    """Stage-1 gate.  Returns ``(probs (n,2) f32, pred (n,) i32, index (n,) i32 [first count valid], count (1,) i32)``,
    all on the device (no synchronisation)."""
    lib = _lib.load()
    logits = _cuda(logits, torch.float32, "gate_compact")
    if logits.dim() != 2 or logits.shape[1] != 2:
        raise RuntimeError("Stage1 output shape unexpected; expected (N,2)")  # ref:310-311
    n = logits.shape[0]
    dev = logits.device
    probs = torch.empty((n, 2), dtype=torch.float32, device=dev)
    pred = torch.empty((n,), dtype=torch.int32, device=dev)
    index = torch.empty((max(n, 1),), dtype=torch.int32, device=dev)
    count = torch.zeros((1,), dtype=torch.int32, device=dev)
    check(lib.zk_gate_compact(logits.data_ptr(), n, float(threshold), -1.0 if min_prob is None else float(min_prob),
                              probs.data_ptr(), pred.data_ptr(), index.data_ptr(), count.data_ptr(),
                              _lib.stream_ptr()), "zk_gate_compact")
    return probs, pred, index, count


# ------------------------------------------------------------------------------------------ model
PFX = "audio_spectrogram_transformer."


class AstModel:
    """Owns a ``zk_model`` (packed bf16 weights on the device) and a reusable activation workspace."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], max_length: int = 1024, num_labels: int = 2,
                 ln_eps: float = 1e-12, num_layers: int = 12, device: Optional[torch.device] = None):
        lib = _lib.load()
        _lib.require_device()
        self._lib = lib
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.max_length, self.num_labels, self.num_layers = max_length, num_labels, num_layers
        keep = []  # keep the fp32 device copies alive until zk_model_create has consumed them

        def p(name: str) -> int:
            if name not in state_dict:
                raise ZkError(f"state dict is missing {name}")
            t = state_dict[name].detach().to(device=self.device, dtype=torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()

        w = _lib.AstWeights()
        w.num_layers, w.max_length, w.num_labels, w.ln_eps = num_layers, max_length, num_labels, ln_eps
        e = PFX + "embeddings."
        w.cls_token, w.dist_token, w.pos_emb = p(e + "cls_token"), p(e + "distillation_token"), p(e + "position_embeddings")
        w.patch_w, w.patch_b = p(e + "patch_embeddings.projection.weight"), p(e + "patch_embeddings.projection.bias")
        for l in range(num_layers):
            q = f"{PFX}encoder.layer.{l}."
            L = w.layer[l]
            L.ln1_w, L.ln1_b = p(q + "layernorm_before.weight"), p(q + "layernorm_before.bias")
            L.q_w, L.q_b = p(q + "attention.attention.query.weight"), p(q + "attention.attention.query.bias")
            L.k_w, L.k_b = p(q + "attention.attention.key.weight"), p(q + "attention.attention.key.bias")
            L.v_w, L.v_b = p(q + "attention.attention.value.weight"), p(q + "attention.attention.value.bias")
            L.o_w, L.o_b = p(q + "attention.output.dense.weight"), p(q + "attention.output.dense.bias")
            L.ln2_w, L.ln2_b = p(q + "layernorm_after.weight"), p(q + "layernorm_after.bias")
            L.fc1_w, L.fc1_b = p(q + "intermediate.dense.weight"), p(q + "intermediate.dense.bias")
            L.fc2_w, L.fc2_b = p(q + "output.dense.weight"), p(q + "output.dense.bias")
        w.final_ln_w, w.final_ln_b = p(PFX + "layernorm.weight"), p(PFX + "layernorm.bias")
        w.head_ln_w, w.head_ln_b = p("classifier.layernorm.weight"), p("classifier.layernorm.bias")
        w.head_w, w.head_b = p("classifier.dense.weight"), p("classifier.dense.bias")
        tokens = 2 + 12 * ((max_length - 16) // 10 + 1)
        pos = state_dict[e + "position_embeddings"]
        if pos.numel() != tokens * HID:
            raise ZkError(f"position table has {pos.numel() // HID} rows, expected {tokens} for max_length {max_length}")
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            check(lib.zk_model_create(C.byref(w), C.byref(h)), "zk_model_create")
        del keep
        self._h = h
        self.tokens = lib.zk_model_num_tokens(h)
        self._ws: Optional[torch.Tensor] = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.zk_model_destroy(h)
            self._h = None

    def workspace_bytes(self, batch: int) -> int:
        return int(self._lib.zk_model_workspace_bytes(self._h, batch))

    def _workspace(self, batch: int) -> torch.Tensor:
        need = self.workspace_bytes(batch)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def forward_features(self, feats: torch.Tensor, return_hidden: bool = False):
        """(B, max_length, 128) normalised features (CUDA fp32) -> logits (B, num_labels) fp32."""
        f = _cuda(feats, torch.float32, "forward_features")
        if f.dim() != 3 or f.shape[1] != self.max_length or f.shape[2] != 128:
            raise ZkError(f"forward_features: expected (B, {self.max_length}, 128), got {tuple(f.shape)}")
        b = f.shape[0]
        logits = torch.empty((b, self.num_labels), dtype=torch.float32, device=f.device)
        hidden = torch.empty((b, self.tokens, HID), dtype=torch.float32, device=f.device) if return_hidden else None
        if b:
            ws = self._workspace(b)
            check(self._lib.zk_model_forward(self._h, f.data_ptr(), b, ws.data_ptr(), ws.numel(), logits.data_ptr(),
                                             hidden.data_ptr() if hidden is not None else None, _lib.stream_ptr()),
                  "zk_model_forward")
        return (logits, hidden) if return_hidden else logits

    def forward_fbank(self, fbank: torch.Tensor, batch: int, mean: float, std: float, window_base: int = 0,
                      window_index: Optional[torch.Tensor] = None, frames_per_hop: int = 50, valid_frames: int = 98,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Fused path: windows are gathered from the compact continuous fbank ``(m,128)`` (un-normalised)."""
        fb = _cuda(fbank, torch.float32, "forward_fbank")
        logits = out if out is not None else torch.empty((batch, self.num_labels), dtype=torch.float32, device=fb.device)
        if batch:
            ws = self._workspace(batch)
            check(self._lib.zk_model_forward_fbank(
                self._h, fb.data_ptr(), fb.shape[0], window_index.data_ptr() if window_index is not None else None,
                window_base, frames_per_hop, valid_frames, float(mean), float(std), batch, ws.data_ptr(), ws.numel(),
                logits.data_ptr(), _lib.stream_ptr()), "zk_model_forward_fbank")
        return logits
