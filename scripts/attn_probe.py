#!/usr/bin/env python
"""Ceiling of the attention softmax instruction stream (scripts/probes/attn_softmax_probe.cu): exp2 / clk / SM for
POLY x key-columns x softmax-warps, nothing else running on the SM.  Usage: python scripts/attn_probe.py [iters]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, "build", "libzk_attn_probe.so"))
lib.zk_attn_softmax_probe.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
grid = 148
clk = torch.zeros((grid, 16, 2), dtype=torch.int64, device="cuda")
lib.zk_mufu_probe.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
clk2 = torch.zeros((grid, 32, 2), dtype=torch.int64, device="cuda")
print("exponential instruction forms: mode (0 f32, 1 f16x2, 2 bf16x2), warps/SM -> instr-lanes / clk / SM")
for mode in (0, 1, 2):
    for warps in (4, 8, 16, 32):
        for rep in range(2):
            clk2.zero_()
            rc = lib.zk_mufu_probe(mode, warps, 4000, clk2.data_ptr(), grid, None)
            torch.cuda.synchronize()
            assert rc == 0, rc
        c = clk2[:, :warps].cpu()
        span = (c[:, :, 1].max(dim=1).values - c[:, :, 0].min(dim=1).values).double().mean().item()
        print(f"  mode {mode} warps {warps:2d}: {warps * 32 * 8 * 4000 / span:6.2f} instr-lanes/clk/SM")
print("poly  w  nsoft   clk/block/warp   exp/clk/SM   us")
for nsoft, w in ((8, 128), (8, 64), (12, 64), (16, 64)):
    for poly in (0, 1, 2):
        for rep in range(2):
            clk.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.zk_attn_softmax_probe(poly, w, nsoft, iters, clk.data_ptr(), grid, None)
            e1.record()
            torch.cuda.synchronize()
            assert rc == 0, rc
        c = clk[:, :nsoft].cpu()
        span = (c[:, :, 1].max(dim=1).values - c[:, :, 0].min(dim=1).values).double()
        per_warp = (c[:, :, 1] - c[:, :, 0]).double().mean().item() / iters
        rate = nsoft * 32 * w * iters / span.mean().item()
        print(f"{poly:4d} {w:4d} {nsoft:5d} {per_warp:14.1f} {rate:12.2f} {e0.elapsed_time(e1) * 1e3:8.1f}")
