"""The fused two-stage sliding-window cascade on one GPU (replaces ref:301-348 / refc:433-568 per recording).

    waveform (host, any rate, any channels)
      -> H2D -> channel mean + polyphase resample to 16 kHz           (zk_resample_*)
      -> ONE continuous Kaldi fbank over the whole recording            (zk_fbank_f32)        SURVEY.md 0.9
      -> Stage 1: batches of windows gathered straight from the compact fbank (frame = 50 w + t), AST forward
      -> decision re-check: windows whose margin is within eps of a threshold are re-run at fp32-class precision
         (zk_band_select -> zk_model_forward_fbank(ZK_PRECISION_RECHECK) -> zk_scatter_rows2), so that ...
      -> softmax + threshold gate + order-preserving compaction         (zk_gate_compact)
         ... decides exactly as the reference's fp32 gate does (ref:312-320)
      -> Stage 2: AST forward on the compacted window index list (same fbank, Stage-2 normalisation) + re-check
      -> one D2H of the per-window scores; summary / aggregation on the host (cascade.py)

The (B,1024,128) feature tensor of the reference (512 KiB per window, 90 % constant padding) is never
materialised.  When the window/hop geometry does not align with the 10 ms frame shift, or the two extractors
differ in more than mean/std, the per-window contract kernels are used instead (still on the GPU).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib, cascade, ops
from ._lib import ZkError
from .fx import ZenkerASTFeatureExtractor
from .model import ZenkerASTForAudioClassification, decision_margins, default_recheck_eps

SAMPLING_RATE = 16000
FRAME_SHIFT = 160


@dataclass
class RecordingResult:
    num_windows: int
    s1_probs: np.ndarray          # (N,2) float32
    s1_preds: np.ndarray          # (N,) int64, (argmax==1) & (p1 >= thr)
    swallow_indices: np.ndarray   # (K,) int64 ascending, windows forwarded to Stage 2
    s2_probs: np.ndarray          # (K,2) float32
    classes: np.ndarray           # (N,) -1 idle / 0 healthy / 1 zenker
    summary: Dict[str, Any] = field(default_factory=dict)
    rechecked_s1: int = 0         # windows re-run at PRECISION_RECHECK before the Stage-1 gate
    rechecked_s2: int = 0         # forwarded windows re-run before the Stage-2 decision

    @property
    def stage2_results(self):
        return [(int(g), self.s2_probs[i]) for i, g in enumerate(self.swallow_indices)]


class TwoStagePipeline:
    def __init__(self, model_s1: ZenkerASTForAudioClassification, fx_s1: ZenkerASTFeatureExtractor,
                 model_s2: ZenkerASTForAudioClassification, fx_s2: ZenkerASTFeatureExtractor, batch_size: int = 128,
                 window_sec: float = 1.0, hop_sec: float = 0.5, stage1_threshold: float = 0.5,
                 stage2_threshold: float = 0.5, stage1_forward_min_prob: Optional[float] = None,
                 stage2_argmax: bool = False, device: Optional[Union[str, torch.device]] = None,
                 recheck_eps: Optional[float] = None, recheck_batch: int = 62):
        if not torch.cuda.is_available():
            raise ZkError("TwoStagePipeline needs a B200; there is no CPU path")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.m1, self.m2 = model_s1.to(self.device), model_s2.to(self.device)
        self.fx1, self.fx2 = fx_s1, fx_s2
        self.batch_size = int(batch_size)
        self.window_sec, self.hop_sec = window_sec, hop_sec
        self.thr1, self.thr2 = float(stage1_threshold), float(stage2_threshold)
        self.min_prob = stage1_forward_min_prob
        self.stage2_argmax = bool(stage2_argmax)
        # decision re-check (module docstring): every threshold the reference applies to a stage's scores is a decision
        # point -- Stage 1: argmax (ref:313, and the bare argmax of summarize, ref:156), p1 >= thr1 (ref:316), the cached
        # script's forward_min_prob (refc:471-478); Stage 2: p_zenker >= thr2 (ref:333) or argmax (refc:512-515)
        self.recheck_eps = float(recheck_eps) if recheck_eps is not None else max(
            default_recheck_eps(self.m1.operand_format), default_recheck_eps(self.m2.operand_format))
        self.recheck_batch = int(recheck_batch)
        thr_s1, thr_s2 = [0.5, self.thr1, self.min_prob], [0.5 if self.stage2_argmax else self.thr2]
        self.margins1, self.margins2 = decision_margins(thr_s1), decision_margins(thr_s2)
        # the drop-in __call__ of the two models (cached flow, cache.py) re-checks against the same decision points
        self.m1.recheck_thresholds, self.m1.recheck_eps = thr_s1, self.recheck_eps
        self.m2.recheck_thresholds, self.m2.recheck_eps = thr_s2, self.recheck_eps
        for fx in (fx_s1, fx_s2):
            if fx.sampling_rate != SAMPLING_RATE:
                raise ZkError("the two-stage path runs at 16 kHz (ref:47)")
        self.win = int(window_sec * SAMPLING_RATE)
        self.hop = int(hop_sec * SAMPLING_RATE)
        d1, d2 = fx_s1.to_dict(), fx_s2.to_dict()
        same_geometry = all(d1[k] == d2[k] for k in d1 if k not in ("mean", "std"))
        self.fused = (same_geometry and self.hop % FRAME_SHIFT == 0 and fx_s1.do_normalize
                      and fx_s1.num_mel_bins == 128 and self.m1.max_length == fx_s1.max_length == self.m2.max_length)
        self.valid_frames = min(ops.FbankPlan.num_frames(self.win), fx_s1.max_length)
        self.plan = fx_s1._get_plan(self.device)
        self._ws: Optional[torch.Tensor] = None  # zk_cascade_run workspace (fbank, logits, model activations)

    # ------------------------------------------------------------------ stages
    def _stage_logits(self, model: ZenkerASTForAudioClassification, fx: ZenkerASTFeatureExtractor, audio: torch.Tensor,
                      fbank: Optional[torch.Tensor], n: int, index: Optional[torch.Tensor],
                      precision: int = _lib.PRECISION_FAST, batch_size: Optional[int] = None) -> torch.Tensor:
        """Logits (n,2) for windows ``index[:n]`` (or 0..n-1 when index is None)."""
        eng = model.engine
        out = torch.empty((n, model.num_labels), dtype=torch.float32, device=self.device)
        B = batch_size or self.batch_size
        for base in range(0, n, B):
            b = min(B, n - base)
            if self.fused:
                eng.forward_fbank(fbank, b, fx.mean, fx.std, window_base=base,
                                  window_index=None if index is None else index[base:base + b],
                                  frames_per_hop=self.hop // FRAME_SHIFT, valid_frames=self.valid_frames,
                                  out=out[base:base + b], precision=precision)
            else:
                if index is None:
                    starts = torch.arange(base, base + b, device=self.device, dtype=torch.int64) * self.hop
                else:
                    starts = index[base:base + b].to(torch.int64) * self.hop
                wins = audio[(starts.unsqueeze(1) + torch.arange(self.win, device=self.device)).reshape(-1)].view(b, self.win)
                feats = fx._get_plan(self.device).fx_contract(wins, fx.mean, fx.std, fx.max_length, fx.do_normalize)
                out[base:base + b] = eng.forward_features(feats, precision=precision)
        return out

    def _recheck(self, model, fx, audio, fbank, logits: torch.Tensor, index: Optional[torch.Tensor], margins) -> int:
        """Re-run the windows of ``logits`` (rows = windows ``index[:n]`` or 0..n-1) that sit within ``recheck_eps`` of a
        decision point at PRECISION_RECHECK and overwrite their rows.  One synchronisation (the count)."""
        n = logits.shape[0]
        if self.recheck_eps <= 0 or not margins or n == 0 or model.num_labels != 2:
            return 0
        pos, window, count = ops.band_select(logits, margins, self.recheck_eps, index)
        r = int(count.item())
        if r:
            hi = self._stage_logits(model, fx, audio, fbank, r, window, precision=_lib.PRECISION_RECHECK,
                                    batch_size=self.recheck_batch)
            ops.scatter_rows2(hi, pos, r, logits)
        return r

    # ------------------------------------------------------------------ re-check band of a new checkpoint
    def measure_fast_margin_error(self, audio: torch.Tensor, max_windows: int = 64) -> Dict[str, float]:
        """``max |(l1 - l0)_fast - (l1 - l0)_recheck|`` per stage over the first ``max_windows`` windows of ``audio`` (CUDA
        float32 mono 16 kHz): the quantity ``recheck_eps`` has to exceed for the gate to decide as the fp32 reference does
        (DESIGN.md 4b).  The default band is sized on the conditioned random-init weights of the tests; a trained
        checkpoint has its own error level, and this is how to read it off a few windows of real audio."""
        with torch.cuda.device(self.device):
            L = int(audio.numel())
            _, _, n = cascade.window_geometry(L, self.window_sec, self.hop_sec)
            if L < self.win:
                audio = torch.cat([audio, torch.zeros(self.win - L, dtype=torch.float32, device=audio.device)])
            n = min(n, int(max_windows))
            audio = audio[:(n - 1) * self.hop + self.win].contiguous()
            fbank = self.plan.fbank(audio) if self.fused else None
            out = {}
            for name, model, fx in (("stage1", self.m1, self.fx1), ("stage2", self.m2, self.fx2)):
                if model.num_labels != 2:
                    continue
                fast = self._stage_logits(model, fx, audio, fbank, n, None)
                slow = self._stage_logits(model, fx, audio, fbank, n, None, precision=_lib.PRECISION_RECHECK,
                                          batch_size=self.recheck_batch)
                out[name] = float(((fast[:, 1] - fast[:, 0]) - (slow[:, 1] - slow[:, 0])).abs().max().item())
            return out

    def calibrate_recheck_eps(self, audio: torch.Tensor, max_windows: int = 64, safety: float = 2.0) -> float:
        """Widen ``recheck_eps`` to ``safety`` x the measured fast-path margin error if that exceeds the current band (it
        is never narrowed); returns the band in force.  The drop-in ``__call__`` of the two models follows."""
        err = self.measure_fast_margin_error(audio, max_windows)
        self.recheck_eps = max(self.recheck_eps, float(safety) * max(err.values(), default=0.0))
        self.m1.recheck_eps = self.m2.recheck_eps = self.recheck_eps
        return self.recheck_eps

    def run_audio16k(self, audio: torch.Tensor, window_range: Optional[Sequence[int]] = None) -> RecordingResult:
        """``audio``: CUDA float32 mono 16 kHz (what ``load_audio`` returns, ref:53-59).

        ``window_range = (w0, w1)`` runs the cascade on windows ``w0 .. w1-1`` of the recording only (a chunk of
        ``dist.shard_window_ranges``): window ``k`` is samples ``[k hop, k hop + win)`` (ref:62-75) and its features are
        frames of those samples alone, so the slice ``[w0 hop, (w1-1) hop + win)`` gives those windows bit for bit;
        indices in the result are relative to ``w0``."""
        if window_range is not None:
            w0, w1 = int(window_range[0]), int(window_range[1])
            _, _, n_all = cascade.window_geometry(int(audio.numel()), self.window_sec, self.hop_sec)
            if not 0 <= w0 < w1 <= n_all:
                raise ZkError(f"window_range {(w0, w1)} outside the recording's {n_all} windows")
            if (w0, w1) != (0, n_all):
                audio = audio[w0 * self.hop:(w1 - 1) * self.hop + self.win]
        with torch.cuda.device(self.device):
            L = int(audio.numel())
            _, _, n = cascade.window_geometry(L, self.window_sec, self.hop_sec)
            if L < self.win:  # ref:70-73: only a too-short recording is zero padded
                audio = torch.cat([audio, torch.zeros(self.win - L, dtype=torch.float32, device=audio.device)])
            if self.fused and self.m1.num_labels == 2 and self.m2.num_labels == 2:
                s1, s1_preds, idx, s2, re1, re2 = self._run_fused(audio.contiguous(), n)
            else:
                s1, s1_preds, idx, s2, re1, re2 = self._run_stepwise(audio, n)
        classes = cascade.stage2_classes(n, idx, s2, self.thr2, self.stage2_argmax)
        summary = cascade.summarize_stage_outputs(s1, idx, s2, self.thr2, self.stage2_argmax)
        return RecordingResult(n, s1, s1_preds, idx, s2, classes, summary, re1, re2)

    def _run_fused(self, audio: torch.Tensor, n: int):
        """One ``zk_cascade_run`` call (include/zk_b200.h section 6): fbank, both stages, re-checks, gate."""
        lib = _lib.load()
        p = _lib.CascadeParams(
            batch_size=self.batch_size, recheck_batch=self.recheck_batch, window_samples=self.win, hop_samples=self.hop,
            mean1=float(self.fx1.mean), std1=float(self.fx1.std), mean2=float(self.fx2.mean), std2=float(self.fx2.std),
            thr1=self.thr1, min_prob=-1.0 if self.min_prob is None else float(self.min_prob), thr2=self.thr2,
            stage2_argmax=1 if self.stage2_argmax else 0, recheck_eps=float(self.recheck_eps))
        h1, h2 = self.m1.engine._h, self.m2.engine._h
        need = int(lib.zk_cascade_workspace_bytes(h1, h2, audio.numel(), C.byref(p)))
        if need == 0:
            raise ZkError(f"zk_cascade_workspace_bytes: {_lib.last_error()}")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        probs1 = torch.empty((n, 2), dtype=torch.float32, device=self.device)
        pred = torch.empty((n,), dtype=torch.int32, device=self.device)
        index = torch.empty((n,), dtype=torch.int32, device=self.device)
        probs2 = torch.empty((n, 2), dtype=torch.float32, device=self.device)
        counts = _lib.CascadeCounts()
        _lib.check(lib.zk_cascade_run(self.plan._h, h1, h2, audio.data_ptr(), audio.numel(), C.byref(p), self._ws.data_ptr(),
                                      self._ws.numel(), probs1.data_ptr(), pred.data_ptr(), index.data_ptr(),
                                      probs2.data_ptr(), C.byref(counts), _lib.stream_ptr()), "zk_cascade_run")
        if counts.num_windows != n:
            raise RuntimeError(f"window count mismatch: {counts.num_windows} != {n}")
        k = int(counts.num_forwarded)
        s1 = probs1.cpu().numpy()  # the D2H copies below wait for the stream
        s1_preds = pred.cpu().numpy().astype(np.int64)
        idx = index[:k].cpu().numpy().astype(np.int64)
        s2 = probs2[:k].cpu().numpy() if k else np.zeros((0, 2), dtype=np.float32)
        return s1, s1_preds, idx, s2, int(counts.rechecked_s1), int(counts.rechecked_s2)

    def _run_stepwise(self, audio: torch.Tensor, n: int):
        """The same cascade driven step by step from Python: the route for window / hop geometries the fused gather does
        not cover (per-window contract kernels), and the cross-check of ``zk_cascade_run`` in the tests."""
        fbank = self.plan.fbank(audio) if self.fused else None
        logits1 = self._stage_logits(self.m1, self.fx1, audio, fbank, n, None)
        if logits1.dim() != 2 or logits1.shape[1] != 2:
            raise RuntimeError("Stage1 output shape unexpected; expected (N,2)")  # ref:310-311
        re1 = self._recheck(self.m1, self.fx1, audio, fbank, logits1, None, self.margins1)
        probs1, pred, index, count = ops.gate_compact(logits1, self.thr1, self.min_prob)
        k = int(count.item())  # Stage 2's batch count depends on it
        re2 = 0
        if k:
            logits2 = self._stage_logits(self.m2, self.fx2, audio, fbank, k, index)
            if logits2.shape[1] != 2:
                raise RuntimeError("Stage2 output shape unexpected; expected (K,2)")  # ref:325-326
            re2 = self._recheck(self.m2, self.fx2, audio, fbank, logits2, index, self.margins2)
            probs2 = ops.softmax2(logits2)
        else:
            probs2 = torch.zeros((0, 2), dtype=torch.float32, device=self.device)
        s1 = probs1.cpu().numpy()
        s1_preds = pred.cpu().numpy().astype(np.int64)
        idx = index[:k].cpu().numpy().astype(np.int64)
        s2 = probs2.cpu().numpy()
        return s1, s1_preds, idx, s2, re1, re2

    def resample_to_device(self, waveform: Union[np.ndarray, torch.Tensor], sample_rate: int) -> torch.Tensor:
        """ref:53-59 (``load_audio`` after the decode): H2D, channel mean, resample -> CUDA float32 mono 16 kHz."""
        w = torch.from_numpy(np.ascontiguousarray(waveform)) if isinstance(waveform, np.ndarray) else waveform
        if w.dtype not in (torch.float32, torch.int16):
            w = w.to(torch.float32)
        with torch.cuda.device(self.device):
            if not w.is_cuda:
                w = (w if w.is_pinned() else w.pin_memory()).to(self.device, non_blocking=True)
            return ops.resample(w, int(sample_rate), SAMPLING_RATE)

    def run_waveform(self, waveform: Union[np.ndarray, torch.Tensor], sample_rate: int,
                     window_range: Optional[Sequence[int]] = None) -> RecordingResult:
        """``waveform``: host (or device) ``(channels, n)`` / ``(n,)`` float32, or ``(n, channels)`` int16 PCM.  The whole
        recording is resampled (the 41-tap filter sees the same neighbours as in an unsplit run) even when only
        ``window_range`` of it is classified."""
        return self.run_audio16k(self.resample_to_device(waveform, sample_rate), window_range)

    def run_patient(self, waveforms: Sequence[Union[np.ndarray, torch.Tensor]], sample_rates: Sequence[int],
                    names: Optional[Sequence[str]] = None) -> Dict[str, Any]:
        """Both files of one patient (ref:301-382): per-file summaries + the patient aggregate."""
        names = list(names) if names is not None else [f"file{i}" for i in range(len(waveforms))]
        per_file: Dict[str, Dict[str, Any]] = {}
        results: List[RecordingResult] = []
        for name, w, sr in zip(names, waveforms, sample_rates):
            r = self.run_waveform(w, sr)
            results.append(r)
            per_file[name] = r.summary
        return {"per_file": per_file, "aggregate": cascade.aggregate_patient(per_file, names), "results": results}
