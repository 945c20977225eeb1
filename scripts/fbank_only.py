#!/usr/bin/env python
"""Runs only the continuous fbank kernel on 1 h of audio a few times (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zenker_audio_detection_b200 import ops  # noqa: E402

plan = ops.FbankPlan()
g = torch.Generator(device="cuda").manual_seed(0)
wave = torch.randn(57_600_000, device="cuda", generator=g) * 0.05
for _ in range(3):
    plan.fbank(wave)
torch.cuda.synchronize()
print("ok")
