#!/usr/bin/env python
"""Our fused attention against torch SDPA (cuDNN / flash) at the bench shape in the regime the cascade runs in: the GPU is
first driven to its power cap with ~150 ms of back-to-back fc1-sized GEMMs (as inside a forward pass), then 20 launches
of one attention implementation are timed with events; the two implementations alternate, so both see the same thermal
and power state.  Prints per-round ms and the SM clock sampled while each implementation runs."""
import os
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zenker_audio_detection_b200 import ops  # noqa: E402

B, T = 128, 1214
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * T, 2304, device="cuda", generator=g)
qkv[:, :1536] *= 2.0
qkv = qkv.half()
q, k, v = (qkv[:, i * 768:(i + 1) * 768].view(B, T, 12, 64).transpose(1, 2) for i in range(3))
a = torch.randn(B * T, 768, device="cuda", generator=g).half()
w = torch.randn(3072, 768, device="cuda", generator=g).half()


class Clock(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.samples, self.tag, self.stop = [], None, False

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append((self.tag, float(out[0]), float(out[1])))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)


def heat(ms=150.0):
    t0 = time.perf_counter()
    while (time.perf_counter() - t0) * 1e3 < ms:
        for _ in range(20):
            torch.matmul(a, w.t())
        torch.cuda.synchronize()


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


ours = lambda: ops.attention(qkv, B, T)  # noqa: E731
sdpa = lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v)  # noqa: E731
for f in (ours, sdpa):
    for _ in range(3):
        f()
clk = Clock()
clk.start()
res = {"ours": [], "sdpa": []}
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for rnd in range(4):
    for name, f in (("ours", ours), ("sdpa", sdpa)) if rnd % 2 == 0 else (("sdpa", sdpa), ("ours", ours)):
        clk.tag = "heat"
        heat()
        clk.tag = name
        res[name].append(timed(f, reps))
clk.stop = True
time.sleep(0.1)
for name in ("ours", "sdpa"):
    cs = [c for t, c, p in clk.samples if t == name]
    ps = [p for t, c, p in clk.samples if t == name]
    print(name, "ms per launch (after heating, %d back-to-back):" % reps, ["%.3f" % x for x in res[name]],
          "SM MHz while running:", (min(cs), sorted(cs)[len(cs) // 2], max(cs)) if cs else None,
          "W:", (min(ps), max(ps)) if ps else None)
