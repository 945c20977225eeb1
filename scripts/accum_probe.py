#!/usr/bin/env python
"""How does the tensor core round when it accumulates?  (design study for the re-check path)

fp16 x fp16 products are exact in fp32, so with fp16-valued inputs the ONLY error of a GEMM is its accumulation.  For a
chain of K/16 tcgen05.mma steps into one TMEM accumulator this prints the error against float64 as
  rel_rms    rms of (got - exact) / rms(exact)
  shrink     mean of (got - exact) * sign(exact) / mean |exact|    (negative = biased towards zero = truncation)
next to the same numbers for torch's fp32 matmul of the same values (FFMA, round to nearest).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zenker_audio_detection_b200 import _lib, ops  # noqa: E402


def stats(got, exact):
    d = got.double() - exact
    return (d.pow(2).mean().sqrt() / exact.pow(2).mean().sqrt()).item(), ((d * exact.sign()).mean() / exact.abs().mean()).item()


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(1)
    M, N = 2048, 768
    for K in (256, 768, 3072):
        for positive in (False, True):
            a = torch.randn(M, K, device="cuda", generator=g)
            w = torch.randn(N, K, device="cuda", generator=g)
            if positive:  # all products positive: the accumulator grows linearly, truncation bias is plain to see
                a, w = a.abs(), w.abs()
            a, w = a.half(), w.half()
            exact = a.double() @ w.double().t()
            x = torch.zeros(M, N, device="cuda")
            ops.gemm(a, w, torch.zeros(N, device="cuda"), _lib.EPI_BIAS_RESID_F32, out=x)
            t32 = a.float() @ w.float().t()
            r1, s1 = stats(x, exact)
            r2, s2 = stats(t32, exact)
            print(f"K={K:5d} {'positive' if positive else 'signed  '}: tcgen05 rel_rms {r1:.3e} shrink {s1:+.3e} | "
                  f"torch fp32 rel_rms {r2:.3e} shrink {s2:+.3e}   ({K // 16} mma steps)")


if __name__ == "__main__":
    main()
