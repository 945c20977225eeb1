#!/usr/bin/env python
"""Times the fused attention kernel alone at the bench shape (CUDA events, L2 flushed between launches); used for
same-box A/B runs of the ZK_ATTN_* switches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zenker_audio_detection_b200 import ops  # noqa: E402

B, T = int(os.environ.get("ZK_BENCH_BATCH", "128")), 1214
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * T, 2304, device="cuda", generator=g)
qkv[:, :1536] *= 2.0
qkv = qkv.to(torch.bfloat16 if os.environ.get("ZK_OPERANDS", "fp16").lower().startswith("b") else torch.float16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5):
    ops.attention(qkv, B, T)
ts = []
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.attention(qkv, B, T)
    b.record()
    b.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
# sustained: back-to-back launches (the regime inside a forward pass: the 1 kW power cap sets the clock)
for _ in range(20):
    ops.attention(qkv, B, T)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(100):
    ops.attention(qkv, B, T)
b.record()
b.synchronize()
sustained = a.elapsed_time(b) / 100
print("sustained ms %.4f" % sustained, end="  ")
print({k: os.environ[k] for k in os.environ if k.startswith("ZK_ATTN")}, "median ms %.4f min %.4f" % (ts[len(ts) // 2], ts[0]))
