#!/usr/bin/env python
"""Times the continuous fbank kernel on cfg3 (1 h of 16 kHz audio, 359 998 frames): same-box A/B runs (ZK_B200_LIB)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zenker_audio_detection_b200 import ops  # noqa: E402

plan = ops.FbankPlan()
g = torch.Generator(device="cuda").manual_seed(3003)
wave = torch.randn(57_600_000, device="cuda", generator=g) * 0.05
m = plan.num_frames(wave.numel())
for _ in range(3):
    plan.fbank(wave)
ts = []
for _ in range(20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    plan.fbank(wave)
    b.record()
    b.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
ms = ts[len(ts) // 2]
gbs = (4.0 * wave.numel() + 512.0 * m) / ms / 1e6
print("fbank cfg3: median %.4f ms  min %.4f  %.0f GB/s  %.3f of 6550 GB/s" % (ms, ts[0], gbs, gbs / 6550.4))
