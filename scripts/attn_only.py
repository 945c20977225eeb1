#!/usr/bin/env python
"""Runs only the fused attention kernel a few times at the bench shape (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zenker_audio_detection_b200 import ops  # noqa: E402

B, T = int(os.environ.get("ZK_BENCH_BATCH", "128")), 1214
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * T, 2304, device="cuda", generator=g)
qkv[:, :1536] *= 2.0
qkv = qkv.to(torch.bfloat16)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    ops.attention(qkv, B, T)
torch.cuda.synchronize()
print("ok")
