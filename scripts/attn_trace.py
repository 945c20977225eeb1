#!/usr/bin/env python
"""Timeline of the attention kernel's per-key-block phases (device clock64), for tuning.
slots per CTA: [0]=smid [1]=t_start [2]=t_softmax_end; per block j at 8+8j: +0 softmax begins waiting for S_j,
+1 S_j ready, +2 P_j published; +4 MMA thread sees P_j, +5 S_{j+1} issued, +6 P_j V_j issued."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zenker_audio_detection_b200 import _lib  # noqa: E402

B, T = 128, 1214
lib = _lib.load()
_lib.require_device()
qkv = torch.randn(B * T, 2304, device="cuda").to(torch.bfloat16)
out = torch.empty(B * T, 768, device="cuda", dtype=torch.bfloat16)
trace = torch.zeros(512 * 128, dtype=torch.int64, device="cuda")
for _ in range(2):
    _lib.check(lib.zk_attention_trace(qkv.data_ptr(), out.data_ptr(), B, T, trace.data_ptr(), _lib.stream_ptr()), "trace")
torch.cuda.synchronize()
t = trace.cpu().numpy().reshape(512, 128)
ncta = min(512, torch.cuda.get_device_properties(0).multi_processor_count)
# slots per block n (first 15 blocks of each CTA, tile A): +0 softmax starts waiting for S_n, +1 S_n ready,
# +3 S_n in registers (buffer handed back), +7 max / rescale done, +2 P_n published; MMA thread: +5 S_n issued, +6 P_n V_n issued
for c in range(2):
    t0 = t[c, 1]
    print(f"CTA {c} on SM {t[c,0]}: first item done at +{t[c,2]-t0}")
    for n in range(15):
        r = t[c, 8 + 8 * n: 16 + 8 * n] - t0
        print(f"   n={n:2d}: S issued {r[5]:7d} | wait {r[0]:7d} S_ready {r[1]:7d} (+{r[1]-r[0]:5d}) loaded {r[3]-r[1]:5d} max {r[7]-r[3]:5d} exp+store {r[2]-r[7]:5d} -> P_pub {r[2]:7d} | PV issued {r[6]:7d} (+{r[6]-r[2]:4d}) | tile B P_pub {r[4]:7d} (A{r[4]-r[2]:+6d})")
d = []
for c in range(ncta):
    for n in range(2, 14):
        r = t[c, 8 + 8 * n: 16 + 8 * n]
        nx = t[c, 8 + 8 * (n + 1): 16 + 8 * (n + 1)]
        d.append((r[1] - r[0], r[3] - r[1], r[7] - r[3], r[2] - r[7], nx[2] - r[2]))
d = np.array(d)
print("median cycles: wait for S %d | load S %d | max %d | exp + P store %d | block period %d" % tuple(np.median(d, axis=0)))
print("mean   cycles: %d %d %d %d %d" % tuple(d.mean(axis=0)))
