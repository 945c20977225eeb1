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
by_sm = {}
for c in range(512):
    by_sm.setdefault(int(t[c, 0]), []).append(c)
print("CTAs per SM among the first 512:", sorted(len(v) for v in by_sm.values())[-5:])
for sm in sorted(by_sm)[:2]:
    ctas = sorted(by_sm[sm], key=lambda c: t[c, 1])[:4]
    t0 = min(t[c, 1] for c in ctas)
    for c in ctas:
        print(f"SM {sm} CTA {c}: start {t[c,1]-t0}, softmax end {t[c,2]-t0} (total {t[c,2]-t[c,1]})")
        for j in range(10):
            r = t[c, 8 + 8 * j: 16 + 8 * j] - t0
            print(f"   j={j}: wait_S {r[0]:7d} S_ready {r[1]:7d} (+{r[1]-r[0]:5d}) P_pub {r[2]:7d} (softmax {r[2]-r[1]:5d}) | mma sees P {r[4]:7d} (+{r[4]-r[2]:4d}) S_next issued {r[5]:7d} PV issued {r[6]:7d}")
d = []
for c in range(512):
    for j in range(1, 9):
        r = t[c, 8 + 8 * j: 16 + 8 * j]
        d.append((r[1] - r[0], r[2] - r[1], r[4] - r[2], t[c, 8 + 8 * (j + 1) + 1] - r[2]))
e = []
for c in range(512):
    for j in range(1, 9):
        r = t[c, 8 + 8 * j: 16 + 8 * j]
        e.append((r[3] - r[1], r[7] - r[3], r[2] - r[7]))
e = np.array(e)
print("softmax split (median): pass 1 (max) %d | rescale check + wait for P V_{j-1} %d | pass 2 (exp, pack, store) %d" % tuple(np.median(e, axis=0)))
d = np.array(d)
print("median cycles: wait for S %d | softmax (S ready -> P published) %d | P published -> MMA thread sees it %d | P published -> next S ready %d"
      % tuple(np.median(d, axis=0)))
print("mean   cycles: %d %d %d %d ; per-block period %d" % (*d.mean(axis=0), (d[:, 1] + d[:, 3]).mean()))
