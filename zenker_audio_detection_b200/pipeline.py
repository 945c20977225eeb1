"""The fused two-stage sliding-window cascade on one GPU (replaces ref:301-348 / refc:433-568 per recording).

    waveform (host, any rate, any channels)
      -> H2D -> channel mean + polyphase resample to 16 kHz           (zk_resample_*)
      -> ONE continuous Kaldi fbank over the whole recording            (zk_fbank_f32)        SURVEY.md 0.9
      -> Stage 1: batches of windows gathered straight from the compact fbank (frame = 50 w + t), AST forward
      -> softmax + threshold gate + order-preserving compaction         (zk_gate_compact)
      -> Stage 2: AST forward on the compacted window index list (same fbank, Stage-2 normalisation)
      -> one D2H of the per-window scores; summary / aggregation on the host (cascade.py)

The (B,1024,128) feature tensor of the reference (512 KiB per window, 90 % constant padding) is never
materialised.  When the window/hop geometry does not align with the 10 ms frame shift, or the two extractors
differ in more than mean/std, the per-window contract kernels are used instead (still on the GPU).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import cascade, ops
from ._lib import ZkError
from .fx import ZenkerASTFeatureExtractor
from .model import ZenkerASTForAudioClassification

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

    @property
    def stage2_results(self):
        return [(int(g), self.s2_probs[i]) for i, g in enumerate(self.swallow_indices)]


class TwoStagePipeline:
    def __init__(self, model_s1: ZenkerASTForAudioClassification, fx_s1: ZenkerASTFeatureExtractor,
                 model_s2: ZenkerASTForAudioClassification, fx_s2: ZenkerASTFeatureExtractor, batch_size: int = 128,
                 window_sec: float = 1.0, hop_sec: float = 0.5, stage1_threshold: float = 0.5,
                 stage2_threshold: float = 0.5, stage1_forward_min_prob: Optional[float] = None,
                 stage2_argmax: bool = False, device: Optional[Union[str, torch.device]] = None):
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
        self.plan = fx_s1._get_plan()

    # ------------------------------------------------------------------ stages
    def _stage_logits(self, model: ZenkerASTForAudioClassification, fx: ZenkerASTFeatureExtractor, audio: torch.Tensor,
                      fbank: Optional[torch.Tensor], n: int, index: Optional[torch.Tensor]) -> torch.Tensor:
        """Logits (n,2) for windows ``index[:n]`` (or 0..n-1 when index is None)."""
        eng = model.engine
        out = torch.empty((n, model.num_labels), dtype=torch.float32, device=self.device)
        B = self.batch_size
        for base in range(0, n, B):
            b = min(B, n - base)
            if self.fused:
                eng.forward_fbank(fbank, b, fx.mean, fx.std, window_base=base,
                                  window_index=None if index is None else index[base:base + b],
                                  frames_per_hop=self.hop // FRAME_SHIFT, valid_frames=self.valid_frames,
                                  out=out[base:base + b])
            else:
                if index is None:
                    starts = torch.arange(base, base + b, device=self.device, dtype=torch.int64) * self.hop
                else:
                    starts = index[base:base + b].to(torch.int64) * self.hop
                wins = audio[(starts.unsqueeze(1) + torch.arange(self.win, device=self.device)).reshape(-1)].view(b, self.win)
                feats = fx._get_plan().fx_contract(wins, fx.mean, fx.std, fx.max_length, fx.do_normalize)
                out[base:base + b] = eng.forward_features(feats)
        return out

    def run_audio16k(self, audio: torch.Tensor) -> RecordingResult:
        """``audio``: CUDA float32 mono 16 kHz (what ``load_audio`` returns, ref:53-59)."""
        with torch.cuda.device(self.device):
            L = int(audio.numel())
            _, _, n = cascade.window_geometry(L, self.window_sec, self.hop_sec)
            if L < self.win:  # ref:70-73: only a too-short recording is zero padded
                audio = torch.cat([audio, torch.zeros(self.win - L, dtype=torch.float32, device=audio.device)])
            fbank = self.plan.fbank(audio) if self.fused else None
            logits1 = self._stage_logits(self.m1, self.fx1, audio, fbank, n, None)
            if logits1.dim() != 2 or logits1.shape[1] != 2:
                raise RuntimeError("Stage1 output shape unexpected; expected (N,2)")  # ref:310-311
            probs1, pred, index, count = ops.gate_compact(logits1, self.thr1, self.min_prob)
            k = int(count.item())  # the only mid-pipeline synchronisation: Stage 2's batch count depends on it
            if k:
                logits2 = self._stage_logits(self.m2, self.fx2, audio, fbank, k, index)
                if logits2.shape[1] != 2:
                    raise RuntimeError("Stage2 output shape unexpected; expected (K,2)")  # ref:325-326
                probs2 = ops.softmax2(logits2)
            else:
                probs2 = torch.zeros((0, 2), dtype=torch.float32, device=self.device)
            s1 = probs1.cpu().numpy()
            s1_preds = pred.cpu().numpy().astype(np.int64)
            idx = index[:k].cpu().numpy().astype(np.int64)
            s2 = probs2.cpu().numpy()
        classes = cascade.stage2_classes(n, idx, s2, self.thr2, self.stage2_argmax)
        summary = cascade.summarize_stage_outputs(s1, idx, s2, self.thr2, self.stage2_argmax)
        return RecordingResult(n, s1, s1_preds, idx, s2, classes, summary)

    def resample_to_device(self, waveform: Union[np.ndarray, torch.Tensor], sample_rate: int) -> torch.Tensor:
        """ref:53-59 (``load_audio`` after the decode): H2D, channel mean, resample -> CUDA float32 mono 16 kHz."""
        w = torch.from_numpy(np.ascontiguousarray(waveform)) if isinstance(waveform, np.ndarray) else waveform
        if w.dtype not in (torch.float32, torch.int16):
            w = w.to(torch.float32)
        with torch.cuda.device(self.device):
            if not w.is_cuda:
                w = (w if w.is_pinned() else w.pin_memory()).to(self.device, non_blocking=True)
            return ops.resample(w, int(sample_rate), SAMPLING_RATE)

    def run_waveform(self, waveform: Union[np.ndarray, torch.Tensor], sample_rate: int) -> RecordingResult:
        """``waveform``: host (or device) ``(channels, n)`` / ``(n,)`` float32, or ``(n, channels)`` int16 PCM."""
        return self.run_audio16k(self.resample_to_device(waveform, sample_rate))

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
