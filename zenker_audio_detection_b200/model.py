"""Drop-in for ``transformers.ASTForAudioClassification`` (inference only) on the two-stage path.

Mirrors the contract the reference scripts use (ref:86-98,108-110): ``from_pretrained(dir,
config=cfg)`` on ``config.json`` + ``model.safetensors`` with the 203 HF tensor names, ``.to(device)``, ``.eval()``,
assignable ``config.label2id / id2label`` and ``__call__(input_values) -> object with .logits`` on the same device.
The forward runs in libzk_b200 (``zk_model_forward``); there is no PyTorch fallback.
"""
from __future__ import annotations

import json
import math
import os
from types import SimpleNamespace
from typing import Any, Dict, Optional

import torch

from . import ops
from ._lib import ZkError

WEIGHTS_SAFE, WEIGHTS_BIN, CONFIG_NAME = "model.safetensors", "pytorch_model.bin", "config.json"
_GEOMETRY = dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, patch_size=16,
                 frequency_stride=10, time_stride=10, num_mel_bins=128)


class SequenceClassifierOutput(dict):
    """``.logits`` / ``["logits"]`` / ``[0]`` access like ``transformers.modeling_outputs.SequenceClassifierOutput``."""

    def __init__(self, logits: torch.Tensor):
        super().__init__(logits=logits)
        self.logits = logits
        self.loss = None
        self.hidden_states = None
        self.attentions = None

    def __getitem__(self, k):
        if isinstance(k, int):
            return (self.logits,)[k]
        return super().__getitem__(k)


def _cfg_get(cfg: Any, name: str, default=None):
    if cfg is None:
        return default
    if isinstance(cfg, dict):
        return cfg.get(name, default)
    return getattr(cfg, name, default)


# Half-width (logit units) of the band around a decision threshold inside which a window is re-run at
# PRECISION_RECHECK.  It has to cover the FAST path's logit error: measured max 2.7e-3, rms ~1e-3 (fp16 operands) and
# 1.9e-2 (bf16) on the conditioned random-init weights; the defaults are ~6 sigma / twice the
# largest error seen.  ZK_RECHECK_EPS overrides.
RECHECK_EPS = {"fp16": 6e-3, "bf16": 4e-2}


def default_recheck_eps(operand_format: str) -> float:
    v = os.environ.get("ZK_RECHECK_EPS")
    return float(v) if v not in (None, "") else RECHECK_EPS["bf16" if operand_format.startswith("b") else "fp16"]


def logit(p: float) -> float:
    """Margin ``l1 - l0`` at which ``softmax([l0, l1])[1] == p``; +-inf at the ends (such a threshold has no band)."""
    p = float(p)
    if p <= 0.0:
        return -math.inf
    if p >= 1.0:
        return math.inf
    return math.log(p / (1.0 - p))


def decision_margins(thresholds) -> list:
    """Finite, de-duplicated decision points (logit units) for a set of probability thresholds on class 1."""
    out = []
    for t in thresholds:
        if t is None:
            continue
        m = logit(t)
        if math.isfinite(m) and all(abs(m - o) > 1e-12 for o in out):
            out.append(m)
    if len(out) > 4:
        raise ZkError("at most 4 distinct decision thresholds per stage are supported by zk_band_select")
    return out


class ZenkerASTForAudioClassification:
    def __init__(self, config: Any, state_dict: Dict[str, torch.Tensor], operand_format: Optional[str] = None):
        for k, v in _GEOMETRY.items():
            got = _cfg_get(config, k, v)
            if got != v:
                raise ZkError(f"config.{k} = {got}: the sm_100a kernels implement the AST-base geometry ({k} = {v}) only")
        act = _cfg_get(config, "hidden_act", "gelu")
        if act != "gelu":
            raise ZkError(f"config.hidden_act = {act!r}; only 'gelu' (erf) is implemented")
        self.config = config if not isinstance(config, dict) else SimpleNamespace(**config)
        self._sd = {k: v.detach().to(torch.float32).cpu() for k, v in state_dict.items()}
        self.num_labels = int(self._sd["classifier.dense.weight"].shape[0])
        self.max_length = int(_cfg_get(config, "max_length", 1024))
        self.ln_eps = float(_cfg_get(config, "layer_norm_eps", 1e-12))
        self.device = torch.device("cpu")
        self._engine: Optional[ops.AstModel] = None
        self.training = False
        self.operand_format = (operand_format or ops.default_operand_format()).lower()
        # Decision re-check inside __call__ (the reference thresholds the returned scores on the host, ref:312-320):
        # windows whose FAST margin is within recheck_eps of a decision point are recomputed at fp32-class precision
        # before the logits are returned.  The decision points default to argmax / p = 0.5 (the scripts' default
        # thresholds, ref:226-227); a caller running other thresholds sets them (or ZK_RECHECK_THRESHOLDS=0.6,0.35).
        env_thr = os.environ.get("ZK_RECHECK_THRESHOLDS")
        self.recheck_thresholds = [float(t) for t in env_thr.split(",")] if env_thr else [0.5]
        self.recheck_eps = default_recheck_eps(self.operand_format)
        self.last_rechecked = 0

    # ------------------------------------------------------------------ loading
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, config: Any = None, operand_format: Optional[str] = None,
                        **kwargs):
        root = pretrained_model_name_or_path
        if config is None:
            path = os.path.join(root, CONFIG_NAME)
            if not os.path.isfile(path):
                raise OSError(f"{path} not found (only local model directories are supported)")
            with open(path, "r", encoding="utf-8") as f:
                config = SimpleNamespace(**json.load(f))
        safe, binf = os.path.join(root, WEIGHTS_SAFE), os.path.join(root, WEIGHTS_BIN)
        if os.path.isfile(safe):
            from safetensors.torch import load_file

            sd = load_file(safe, device="cpu")
        elif os.path.isfile(binf):
            sd = torch.load(binf, map_location="cpu", weights_only=True)
        else:
            raise OSError(f"no {WEIGHTS_SAFE} or {WEIGHTS_BIN} under {root}")
        return cls(config, sd, operand_format=operand_format)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return dict(self._sd)

    # ------------------------------------------------------------------ nn.Module-like surface
    def to(self, device=None, *args, **kwargs):
        if device is None:
            return self
        device = torch.device(device)
        if device.type != "cuda":
            raise ZkError("ZenkerASTForAudioClassification runs on a B200 only; there is no CPU path")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if self._engine is None or self.device != device:
            with torch.cuda.device(device):
                self._engine = ops.AstModel(self._sd, self.max_length, self.num_labels, self.ln_eps, device=device,
                                            operand_format=self.operand_format)
            self.device = device
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", device if device is not None else torch.cuda.current_device()))

    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        if mode:
            raise ZkError("inference only: training is out of scope for the zenker-b200 path")
        return self

    def parameters(self):
        return iter(self._sd.values())

    @property
    def engine(self) -> ops.AstModel:
        if self._engine is None:
            self.to("cuda")
        return self._engine

    def forward(self, input_values: Optional[torch.Tensor] = None, labels=None, **kwargs) -> SequenceClassifierOutput:
        if input_values is None:
            raise ValueError("You have to specify input_values")  # HF:modeling...:371-372
        if labels is not None:
            raise ZkError("inference only: `labels` / loss are out of scope")
        eng = self.engine
        x = input_values
        if not x.is_cuda:
            x = x.to(self.device)
        if x.dtype != torch.float32:
            x = x.float()
        with torch.cuda.device(self.device):
            logits = eng.forward_features(x)
            self.last_rechecked = 0
            if self.num_labels == 2 and self.recheck_eps > 0:
                self.last_rechecked = eng.recheck_features(x, logits, decision_margins(self.recheck_thresholds),
                                                           self.recheck_eps)
        return SequenceClassifierOutput(logits)

    __call__ = forward
