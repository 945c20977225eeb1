#!/usr/bin/env python
"""Ceiling of the attention softmax instruction stream (scripts/probes/attn_softmax_probe.cu): exp2 / clk / SM for
POLY x key-columns x softmax-warps, nothing else running on the SM.  Usage: python scripts/attn_probe.py [iters]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, "build", "libzk_attn_probe.so"))
lib.zk_attn_softmax_probe.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
only = tuple(int(v) for v in sys.argv[2].split(",")) if len(sys.argv) > 2 else None  # poly,w,nsoft,mma,idle: one config (ncu)
grid = 148
clk = torch.zeros((grid * 32 + grid,), dtype=torch.int64, device="cuda")
lib.zk_mufu_probe.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
clk2 = torch.zeros((grid, 32, 2), dtype=torch.int64, device="cuda")
if only:
    lib.zk_attn_softmax_probe(only[0], only[1], only[2], iters, clk.data_ptr(), grid, 0, only[3], only[4], None)
    torch.cuda.synchronize()
    sys.exit(0)
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
print("poly  w  nsoft  mma idle_ns  clk/block/warp   exp/clk/SM   tensor-busy")
for nsoft, w, mma, idle in ((8, 128, 0, 0), (8, 128, 1, 0), (8, 128, 2, 0), (8, 128, 3, 0), (8, 128, 3, 500), (8, 128, 3, 1000),
                            (8, 128, 3, 2000), (8, 64, 0, 0), (12, 64, 0, 0)):
    for poly in (0, 1):
        for rep in range(2):
            clk.zero_()
            rc = lib.zk_attn_softmax_probe(poly, w, nsoft, iters, clk.data_ptr(), grid, 0, mma, idle, None)
            torch.cuda.synchronize()
            assert rc == 0, rc
        c = clk[:grid * 32].view(grid, 16, 2)[:, :nsoft].cpu()
        span = (c[:, :, 1].max(dim=1).values - c[:, :, 0].min(dim=1).values).double()
        per_warp = (c[:, :, 1] - c[:, :, 0]).double().mean().item() / iters
        rate = nsoft * 32 * w * iters / span.mean().item()
        batches = clk[grid * 32:].double().mean().item()
        busy = batches * ((256 if mma & 1 else 0) + (256 if mma & 2 else 0)) / span.mean().item()
        print(f"{poly:4d} {w:4d} {nsoft:5d} {mma:4d} {idle:6d} {per_warp:14.1f} {rate:12.2f} {busy:10.2f}")
