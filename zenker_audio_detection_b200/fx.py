"""Drop-in for ``transformers.ASTFeatureExtractor`` on the two-stage inference path.

Mirrors HF:feature_extraction_audio_spectrogram_transformer.py:40-232 as far as the reference scripts touch it
(SURVEY.md section 8b): ``from_pretrained`` / ``save_pretrained`` on ``preprocessor_config.json``, mutable
``mean/std/max_length/...`` attributes, ``to_dict``, ``model_input_names`` and ``__call__`` with the same argument
meaning and error behaviour.  The arithmetic runs in the sm_100a fbank kernel (``zk_fx_contract_f32``); with
``return_tensors="pt"`` the features stay on the GPU (callers do ``.to(DEVICE)``, ref:109).
"""
from __future__ import annotations

import json
import os
from typing import Any, Dict, List, Optional, Union

import numpy as np
import torch

from . import ops
from ._lib import ZkError

CONFIG_NAME = "preprocessor_config.json"


class BatchFeature(dict):
    """Minimal stand-in for ``transformers.BatchFeature``: a dict with attribute access and ``.to``."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def to(self, *args, **kwargs):
        for k, v in list(self.items()):
            if isinstance(v, torch.Tensor):
                self[k] = v.to(*args, **kwargs)
        return self


class ZenkerASTFeatureExtractor:
    model_input_names = ["input_values"]

    def __init__(self, feature_size: int = 1, sampling_rate: int = 16000, num_mel_bins: int = 128,
                 max_length: int = 1024, padding_value: float = 0.0, do_normalize: bool = True,
                 mean: float = -4.2677393, std: float = 4.5689974, return_attention_mask: bool = False,
                 padding_side: str = "right", **kwargs):
        self.feature_size = feature_size
        self.sampling_rate = sampling_rate
        self.padding_value = padding_value
        self.padding_side = padding_side
        self.return_attention_mask = return_attention_mask
        self.num_mel_bins = num_mel_bins
        self.max_length = max_length
        self.do_normalize = do_normalize
        self.mean = mean
        self.std = std
        self._plan: dict = {}  # device index -> ops.FbankPlan

    # ------------------------------------------------------------------ (de)serialisation
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, **kwargs) -> "ZenkerASTFeatureExtractor":
        path = os.path.join(pretrained_model_name_or_path, CONFIG_NAME)
        if not os.path.isfile(path):
            raise OSError(f"Can't load feature extractor for '{pretrained_model_name_or_path}': {CONFIG_NAME} not found "
                          "(only local directories are supported; there is no hub access)")
        with open(path, "r", encoding="utf-8") as f:
            cfg = json.load(f)
        cfg.pop("feature_extractor_type", None)
        cfg.pop("processor_class", None)
        cfg.update(kwargs)
        return cls(**cfg)

    def to_dict(self) -> Dict[str, Any]:
        return {
            "feature_size": self.feature_size, "sampling_rate": self.sampling_rate,
            "padding_value": self.padding_value, "padding_side": self.padding_side,
            "return_attention_mask": self.return_attention_mask, "num_mel_bins": self.num_mel_bins,
            "max_length": self.max_length, "do_normalize": self.do_normalize, "mean": self.mean, "std": self.std,
            "feature_extractor_type": "ASTFeatureExtractor",
        }

    def to_json_string(self) -> str:
        return json.dumps(self.to_dict(), indent=2, sort_keys=True) + "\n"

    def save_pretrained(self, save_directory: str, **kwargs) -> List[str]:
        os.makedirs(save_directory, exist_ok=True)
        path = os.path.join(save_directory, CONFIG_NAME)
        with open(path, "w", encoding="utf-8") as f:
            f.write(self.to_json_string())
        return [path]

    def __repr__(self) -> str:
        return f"{self.__class__.__name__} {self.to_json_string()}"

    # ------------------------------------------------------------------ compute
    def _get_plan(self, device: Optional[torch.device] = None) -> ops.FbankPlan:
        """The fbank tables for ``device`` (default: the current CUDA device); one plan per device, because the tables
        live in that device's memory (zk_fbank_f32 rejects a plan from another device)."""
        idx = torch.device(device).index if device is not None else None
        if idx is None:
            idx = torch.cuda.current_device()
        if not isinstance(self._plan, dict):
            self._plan = {}
        plan = self._plan.get(idx)
        if plan is None or plan.num_mel_bins != self.num_mel_bins:
            with torch.cuda.device(idx):
                plan = ops.FbankPlan("hanning", self.num_mel_bins)  # HF:...:116-121 window_type="hanning"
            self._plan[idx] = plan
        return plan

    def features_cuda(self, raw: List[np.ndarray]) -> torch.Tensor:
        """(B, max_length, num_mel_bins) float32 CUDA tensor for a list of mono float32 waveforms."""
        lens = {int(w.shape[0]) for w in raw}
        plan = self._get_plan()
        if len(lens) == 1:
            host = torch.from_numpy(np.ascontiguousarray(np.stack(raw, axis=0)))
            dev = host.pin_memory().cuda(non_blocking=True) if host.numel() > (1 << 16) else host.cuda()
            return plan.fx_contract(dev, self.mean, self.std, self.max_length, self.do_normalize)
        out = []
        for w in raw:  # ragged batch: one launch per distinct waveform (the reference loops per waveform too)
            dev = torch.from_numpy(np.ascontiguousarray(w)).cuda().unsqueeze(0)
            out.append(plan.fx_contract(dev, self.mean, self.std, self.max_length, self.do_normalize))
        return torch.cat(out, dim=0)

    def __call__(self, raw_speech: Union[np.ndarray, List[float], List[np.ndarray], List[List[float]]],
                 sampling_rate: Optional[int] = None, return_tensors: Optional[str] = None, **kwargs) -> BatchFeature:
        if sampling_rate is not None and sampling_rate != self.sampling_rate:
            raise ValueError(
                f"The model corresponding to this feature extractor: {self} was trained using a sampling rate of"
                f" {self.sampling_rate}. Please make sure that the provided `raw_speech` input was sampled with"
                f" {self.sampling_rate} and not {sampling_rate}.")
        if isinstance(raw_speech, torch.Tensor):
            raw_speech = raw_speech.detach().cpu().numpy()
        is_batched_numpy = isinstance(raw_speech, np.ndarray) and len(raw_speech.shape) > 1
        if is_batched_numpy and len(raw_speech.shape) > 2:
            raise ValueError(f"Only mono-channel audio is supported for input to {self}")
        is_batched = is_batched_numpy or (
            isinstance(raw_speech, (list, tuple)) and len(raw_speech) > 0
            and isinstance(raw_speech[0], (np.ndarray, tuple, list)))
        if is_batched:
            raw = [np.asarray(s, dtype=np.float32) for s in raw_speech]
        else:
            raw = [np.asarray(raw_speech, dtype=np.float32)]  # always return a batch (HF:...:211-212)
        if any(w.ndim != 1 for w in raw):
            raise ValueError(f"Only mono-channel audio is supported for input to {self}")
        feats = self.features_cuda(raw)
        tt = return_tensors.value if hasattr(return_tensors, "value") else return_tensors
        if tt is None:
            arr = feats.cpu().numpy()
            return BatchFeature({"input_values": [a for a in arr]})
        if tt == "pt":
            return BatchFeature({"input_values": feats})
        if tt == "np":
            return BatchFeature({"input_values": feats.cpu().numpy()})
        raise ValueError(f"return_tensors={return_tensors!r} is not supported (use None, 'pt' or 'np')")
