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
_DT = {_lib.FMT_BF16: torch.bfloat16, _lib.FMT_F16: torch.float16}


def _fmt_of(t: torch.Tensor, name: str) -> int:
    if t.dtype == torch.bfloat16:
        return _lib.FMT_BF16
    if t.dtype == torch.float16:
        return _lib.FMT_F16
    raise ZkError(f"{name}: expected a bfloat16 or float16 tensor, got {t.dtype}")


def split_f16(x: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """CUDA fp32 ``(rows, cols)`` -> fp16 ``(rows, 2*cols)`` = hi | lo planes of ``x*scale`` (split-operand layout)."""
    lib = _lib.load()
    x = _cuda(x, torch.float32, "split_f16")
    rows, cols = x.shape
    out = torch.empty((rows, 2 * cols), dtype=torch.float16, device=x.device)
    check(lib.zk_f32_to_16(x.data_ptr(), out.data_ptr(), rows, cols, _lib.FMT_F16, 2, float(scale), _lib.stream_ptr()),
          "zk_f32_to_16")
    return out


def gemm(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, epilogue: int, out: Optional[torch.Tensor] = None,
         aux: Optional[torch.Tensor] = None, aux_rows: int = 0, products: int = 1, acc_scale: float = 1.0) -> torch.Tensor:
    """``epilogue(a @ w.T * acc_scale + bias)`` on the tcgen05 GEMM.  ``a`` (M, planes*K), ``w`` (N, planes*K) in bf16 or
    fp16; ``products == 3``: fp16 hi | lo planes (planes = 2) and the three-product sum."""
    lib = _lib.load()
    fmt = _fmt_of(a, "gemm a")
    a = _cuda(a, _DT[fmt], "gemm a")
    w = _cuda(w, _DT[fmt], "gemm w")
    bias = _cuda(bias, torch.float32, "gemm bias")
    planes = 2 if products == 3 else 1
    M, K = a.shape[0], a.shape[1] // planes
    N = w.shape[0]
    if w.shape[1] != planes * K:
        raise ZkError(f"gemm: a is (M, {a.shape[1]}) but w is (N, {w.shape[1]})")
    if out is None:
        if epilogue in (_lib.EPI_BIAS_BF16, _lib.EPI_BIAS_GELU_BF16):
            out = torch.empty((M, N), dtype=_DT[fmt], device=a.device)
        elif epilogue in (_lib.EPI_BIAS_SPLIT, _lib.EPI_BIAS_GELU_SPLIT):
            out = torch.empty((M, 2 * N), dtype=torch.float16, device=a.device)
        else:
            raise ZkError("gemm: the fp32 epilogues accumulate into / scatter to a caller-provided `out`")
    check(lib.zk_gemm16(a.data_ptr(), 0, w.data_ptr(), 0, bias.data_ptr(), out.data_ptr(), 0, M, N, K, epilogue, fmt,
                        products, float(acc_scale), aux.data_ptr() if aux is not None else None, aux_rows,
                        _lib.stream_ptr()), "zk_gemm16")
    return out


def layernorm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float, dtype: torch.dtype = torch.bfloat16,
              planes: int = 1) -> torch.Tensor:
    lib = _lib.load()
    x = _cuda(x, torch.float32, "layernorm x")
    rows, cols = x.shape
    fmt = _lib.FMT_F16 if dtype == torch.float16 else _lib.FMT_BF16
    out = torch.empty((rows, planes * cols), dtype=dtype, device=x.device)
    check(lib.zk_layernorm16(x.data_ptr(), _cuda(w, torch.float32, "w").data_ptr(),
                             _cuda(b, torch.float32, "b").data_ptr(), eps, out.data_ptr(), rows, cols, fmt, planes,
                             _lib.stream_ptr()), "zk_layernorm16")
    return out


def attention(qkv: torch.Tensor, batch: int, tokens: int) -> torch.Tensor:
    lib = _lib.load()
    fmt = _fmt_of(qkv, "attention qkv")
    qkv = _cuda(qkv, _DT[fmt], "attention qkv")
    if qkv.shape != (batch * tokens, 3 * HID):
        raise ZkError(f"attention: qkv must be ({batch * tokens}, {3 * HID}), got {tuple(qkv.shape)}")
    out = torch.empty((batch * tokens, HID), dtype=_DT[fmt], device=qkv.device)
    check(lib.zk_attention16(qkv.data_ptr(), out.data_ptr(), batch, tokens, fmt, _lib.stream_ptr()), "zk_attention16")
    return out


def attention_split(qkv: torch.Tensor, batch: int, tokens: int) -> torch.Tensor:
    """Re-check precision: ``qkv`` fp16 (batch*tokens, 2*2304) hi | lo planes -> fp16 (batch*tokens, 2*768) hi | lo."""
    lib = _lib.load()
    qkv = _cuda(qkv, torch.float16, "attention_split qkv")
    if qkv.shape != (batch * tokens, 6 * HID):
        raise ZkError(f"attention_split: qkv must be ({batch * tokens}, {6 * HID}), got {tuple(qkv.shape)}")
    out = torch.empty((batch * tokens, 2 * HID), dtype=torch.float16, device=qkv.device)
    check(lib.zk_attention_split(qkv.data_ptr(), out.data_ptr(), batch, tokens, _lib.stream_ptr()), "zk_attention_split")
    return out


class FeatureStats:
    """Running mean / std of un-normalised features, as utils/compute_ast_normalization_stats.py:55-95 computes them
    (float64 sums over every element of the zero-padded ``(B, max_length, 128)`` tensors; unbiased std)."""

    def __init__(self, device: Optional[Union[str, torch.device]] = None):
        self.acc = torch.zeros(2, dtype=torch.float64, device=device or torch.device("cuda", torch.cuda.current_device()))
        self.count = 0

    def update_from_waveforms(self, plan: "FbankPlan", windows: torch.Tensor, max_length: int,
                              return_features: bool = False) -> Optional[torch.Tensor]:
        """``windows``: CUDA float32 ``(B, samples)`` equal-length waveforms.  Their un-normalised features are summed
        INSIDE the feature kernel (``zk_fx_stats_f32``); nothing is written unless ``return_features``."""
        lib = _lib.load()
        w = _cuda(windows, torch.float32, "FeatureStats.update_from_waveforms")
        if w.dim() != 2:
            raise ZkError("update_from_waveforms: expected (batch, samples)")
        b, n = w.shape
        out = torch.empty((b, max_length, 128), dtype=torch.float32, device=w.device) if return_features else None
        if b:
            check(lib.zk_fx_stats_f32(plan._h, w.data_ptr(), b, n, n, max_length, out.data_ptr() if out is not None else None,
                                      self.acc.data_ptr(), _lib.stream_ptr()), "zk_fx_stats_f32")
        self.count += b * max_length * 128
        return out

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


def band_select(logits: torch.Tensor, margins, eps: float, src_window: Optional[torch.Tensor] = None):
    """Rows whose margin ``l1 - l0`` is within ``eps`` of one of ``margins`` (<= 4 decision points, logit units).
    Returns ``(pos (n,) i32, window (n,) i32, count (1,) i32)`` on the device; the first ``count`` entries are valid,
    ascending.  ``window[j] = src_window[pos[j]]`` when a compacted index list is given, else ``pos[j]``."""
    lib = _lib.load()
    logits = _cuda(logits, torch.float32, "band_select")
    n = logits.shape[0]
    dev = logits.device
    pos = torch.empty((max(n, 1),), dtype=torch.int32, device=dev)
    window = torch.empty((max(n, 1),), dtype=torch.int32, device=dev)
    count = torch.zeros((1,), dtype=torch.int32, device=dev)
    mg = (C.c_float * 4)(*([float(m) for m in margins] + [0.0] * (4 - len(margins))))
    check(lib.zk_band_select(logits.data_ptr(), n, mg, len(margins), float(eps),
                             src_window.data_ptr() if src_window is not None else None, pos.data_ptr(),
                             window.data_ptr(), count.data_ptr(), _lib.stream_ptr()), "zk_band_select")
    return pos, window, count


def scatter_rows2(src: torch.Tensor, pos: torch.Tensor, count: int, dst: torch.Tensor) -> None:
    """``dst[pos[j]] = src[j]`` for ``j < count`` (rows of two fp32 values)."""
    lib = _lib.load()
    check(lib.zk_scatter_rows2(_cuda(src, torch.float32, "scatter src").data_ptr(), pos.data_ptr(), int(count),
                               dst.data_ptr(), _lib.stream_ptr()), "zk_scatter_rows2")


# ------------------------------------------------------------------------------------------ model
PFX = "audio_spectrogram_transformer."
_FORMATS = {"fp16": _lib.FMT_F16, "f16": _lib.FMT_F16, "float16": _lib.FMT_F16, "bf16": _lib.FMT_BF16,
            "bfloat16": _lib.FMT_BF16}


def default_operand_format() -> str:
    """fp16 unless ``ZK_OPERANDS=bf16``: tcgen05 kind::f16 runs both at the same rate, fp16 keeps 3 more bits."""
    import os

    v = os.environ.get("ZK_OPERANDS", "fp16").lower()
    if v not in _FORMATS:
        raise ZkError(f"ZK_OPERANDS={v!r}: expected fp16 or bf16")
    return v


class AstModel:
    """Owns a ``zk_model`` (packed 16-bit weights on the device) and a reusable activation workspace."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], max_length: int = 1024, num_labels: int = 2,
                 ln_eps: float = 1e-12, num_layers: int = 12, device: Optional[torch.device] = None,
                 operand_format: Optional[str] = None):
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
        self.operand_format = (operand_format or default_operand_format()).lower()
        if self.operand_format not in _FORMATS:
            raise ZkError(f"operand_format {operand_format!r}: expected 'fp16' or 'bf16'")
        w.operand_format = _FORMATS[self.operand_format]
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

    def workspace_bytes(self, batch: int, precision: int = _lib.PRECISION_FAST) -> int:
        return int(self._lib.zk_model_workspace_bytes(self._h, batch, precision))

    def _workspace(self, batch: int, precision: int = _lib.PRECISION_FAST) -> torch.Tensor:
        need = self.workspace_bytes(batch, precision)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def forward_features(self, feats: torch.Tensor, return_hidden: bool = False, precision: int = _lib.PRECISION_FAST,
                         row_index: Optional[torch.Tensor] = None, batch: Optional[int] = None):
        """(B, max_length, 128) normalised features (CUDA fp32) -> logits (B, num_labels) fp32.  With ``row_index``
        (CUDA int32) window ``i`` of the ``batch`` reads ``feats[row_index[i]]``."""
        f = _cuda(feats, torch.float32, "forward_features")
        if f.dim() != 3 or f.shape[1] != self.max_length or f.shape[2] != 128:
            raise ZkError(f"forward_features: expected (B, {self.max_length}, 128), got {tuple(f.shape)}")
        b = f.shape[0] if row_index is None else int(batch if batch is not None else row_index.numel())
        logits = torch.empty((b, self.num_labels), dtype=torch.float32, device=f.device)
        hidden = torch.empty((b, self.tokens, HID), dtype=torch.float32, device=f.device) if return_hidden else None
        if b:
            ws = self._workspace(b, precision)
            check(self._lib.zk_model_forward(self._h, f.data_ptr(), row_index.data_ptr() if row_index is not None else None,
                                             b, precision, ws.data_ptr(), ws.numel(), logits.data_ptr(),
                                             hidden.data_ptr() if hidden is not None else None, _lib.stream_ptr()),
                  "zk_model_forward")
        return (logits, hidden) if return_hidden else logits

    def recheck_features(self, feats: torch.Tensor, logits: torch.Tensor, margins, eps: float, chunk: int = 62) -> int:
        """Decision re-check on the contract path: rows of ``logits`` whose margin is within ``eps`` of a decision point
        are recomputed at ``PRECISION_RECHECK`` from the same feature rows and overwritten in place.  Returns how many."""
        if eps <= 0 or not len(margins) or logits.shape[0] == 0:
            return 0
        pos, _, count = band_select(logits, margins, eps)
        r = int(count.item())
        for base in range(0, r, chunk):
            n = min(chunk, r - base)
            hi = self.forward_features(feats, precision=_lib.PRECISION_RECHECK, row_index=pos[base:base + n], batch=n)
            scatter_rows2(hi, pos[base:base + n], n, logits)
        return r

    def forward_fbank(self, fbank: torch.Tensor, batch: int, mean: float, std: float, window_base: int = 0,
                      window_index: Optional[torch.Tensor] = None, frames_per_hop: int = 50, valid_frames: int = 98,
                      out: Optional[torch.Tensor] = None, precision: int = _lib.PRECISION_FAST) -> torch.Tensor:
        """Fused path: windows are gathered from the compact continuous fbank ``(m,128)`` (un-normalised)."""
        fb = _cuda(fbank, torch.float32, "forward_fbank")
        logits = out if out is not None else torch.empty((batch, self.num_labels), dtype=torch.float32, device=fb.device)
        if batch:
            ws = self._workspace(batch, precision)
            check(self._lib.zk_model_forward_fbank(
                self._h, fb.data_ptr(), fb.shape[0], window_index.data_ptr() if window_index is not None else None,
                window_base, frames_per_hop, valid_frames, float(mean), float(std), batch, precision, ws.data_ptr(),
                ws.numel(), logits.data_ptr(), _lib.stream_ptr()), "zk_model_forward_fbank")
        return logits
