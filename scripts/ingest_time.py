#!/usr/bin/env python
"""Host side of the ingest (ref:53-59 after the decode): time from "samples in host memory" to "16 kHz mono on the GPU"
for a 10-minute stereo PCM16 recording at 48 kHz (115 MB), (a) from pageable memory, as ``wavio.read`` delivers by
default (the pipeline stages it through a pinned copy on the calling thread), (b) from page-locked memory, as
``batch.read_pinned`` delivers.  The difference is main-thread time during which the GPU has nothing to do."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zenker_audio_detection_b200 import ops  # noqa: E402


def to_device_and_resample(w: torch.Tensor) -> torch.Tensor:
    w = (w if w.is_pinned() else w.pin_memory()).to("cuda", non_blocking=True)
    return ops.resample(w, 48000, 16000)


def main():
    torch.cuda.set_device(0)
    n = 28_800_000
    pageable = torch.from_numpy(np.random.default_rng(0).integers(-3000, 3000, (n, 2), dtype=np.int16))
    pinned = torch.empty((n, 2), dtype=torch.int16, pin_memory=True)
    pinned.copy_(pageable)
    out = {}
    for name, w in (("pageable", pageable), ("pinned", pinned), ("pageable_again", pageable), ("pinned_again", pinned)):
        ts = []
        for _ in range(6):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            y = to_device_and_resample(w)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        out[name + "_ms"] = round(float(np.median(ts[1:])), 2)
    out["samples_16k"] = int(y.numel())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
