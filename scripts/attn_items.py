#!/usr/bin/env python
"""Per-item timeline of the attention kernel (ZK_ATTN_TRACE_ITEMS=1): how long each CTA spends per work item over the
whole launch, and how far tile B runs behind tile A -- the kernel's run-to-run spread comes from that phase."""
import os
import sys

import numpy as np
import torch

os.environ["ZK_ATTN_TRACE_ITEMS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zenker_audio_detection_b200 import _lib  # noqa: E402

B, T = 128, 1214
lib = _lib.load()
_lib.require_device()
qkv = torch.randn(B * T, 2304, device="cuda")
qkv[:, :1536] *= 2.0
qkv = qkv.to(torch.bfloat16)
out = torch.empty(B * T, 768, device="cuda", dtype=torch.bfloat16)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    trace = torch.zeros(512 * 128, dtype=torch.int64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.zk_attention_trace(qkv.data_ptr(), out.data_ptr(), B, T, trace.data_ptr(), _lib.stream_ptr()), "trace")
    e1.record()
    torch.cuda.synchronize()
    t = trace.cpu().numpy().reshape(512, 128)[:148]
    a, b = t[:, 8:8 + 51], t[:, 68:68 + 51]
    dur = np.diff(a, axis=1)            # item durations of tile A per CTA
    lag = (b - a)[:, 1:]                # tile B end - tile A end per item
    total = (np.maximum(a[:, 50], b[:, 50]) - t[:, 1])
    print(f"launch {rep}: {e0.elapsed_time(e1):.3f} ms; CTA span median {np.median(total):.0f} max {total.max():.0f} clk; "
          f"item median {np.median(dur):.0f} p10 {np.percentile(dur,10):.0f} p90 {np.percentile(dur,90):.0f}; "
          f"B-A lag median {np.median(lag):.0f} p10 {np.percentile(lag,10):.0f} p90 {np.percentile(lag,90):.0f}")
    # relation between lag and item duration
    l, d = lag[:, :-1].ravel(), dur[:, 1:].ravel()
    for lo, hi in ((-1e9, -2000), (-2000, -1000), (-1000, -300), (-300, 300), (300, 1000), (1000, 2000), (2000, 1e9)):
        m = (l >= lo) & (l < hi)
        if m.sum() > 20:
            print(f"    lag in [{lo:.0f},{hi:.0f}): {m.sum():5d} items, median duration {np.median(d[m]):.0f}")
