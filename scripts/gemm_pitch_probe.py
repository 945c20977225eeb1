#!/usr/bin/env python
"""Why are the fast GEMMs ~7 % slower in fp16 with plane-interleaved weights than round 1's bf16?  Separates the two
suspects on the fc1 / QKV shapes (M = 128 x 1214): operand format (fp16 multipliers draw more power than bf16 under the
1 kW cap) and weight row pitch (hi | lo planes interleaved per row: the W tile rows are 2K apart)."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zenker_audio_detection_b200 import _lib  # noqa: E402


def timeit(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    lib = _lib.load()
    M = 128 * 1214
    out = {}
    g = torch.Generator(device="cuda").manual_seed(0)
    for name, N, K, epi in (("fc1", 3072, 768, _lib.EPI_BIAS_GELU_BF16), ("qkv", 2304, 768, _lib.EPI_BIAS_BF16),
                            ("fc2", 768, 3072, _lib.EPI_BIAS_RESID_F32)):
        a32 = torch.randn(M, K, device="cuda", generator=g) * 0.5
        w32 = torch.randn(N, K, device="cuda", generator=g) * 0.02
        b = torch.randn(N, device="cuda", generator=g) * 0.1
        x = torch.randn(M, N, device="cuda", generator=g) if epi == _lib.EPI_BIAS_RESID_F32 else None
        for dt, fmt in ((torch.bfloat16, _lib.FMT_BF16), (torch.float16, _lib.FMT_F16)):
            a = a32.to(dt)
            o = x if x is not None else torch.empty(M, N, device="cuda", dtype=dt)
            for pitch in (1, 2):
                w = torch.zeros(N, pitch * K, device="cuda", dtype=dt)
                w[:, :K] = w32.to(dt)
                s = torch.cuda.current_stream().cuda_stream
                fn = lambda: _lib.check(lib.zk_gemm16(a.data_ptr(), 0, w.data_ptr(), pitch * K, b.data_ptr(), o.data_ptr(), 0, M, N,
                                                      K, epi, fmt, 1, 1.0, None, 0, s), "zk_gemm16")
                out[f"{name}_{str(dt).split('.')[-1]}_pitch{pitch}K"] = round(timeit(fn), 4)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
