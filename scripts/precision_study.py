"""CPU emulation of the operand formats of the AST forward (design study for the re-check path).

Every matmul operand (LayerNorm output, q/k/v, softmax probabilities, attention output, GELU output, weights) is
rounded to the format under study; products are accumulated in float64 so that only the OPERAND representation
error is visible.  Compared against the float64 forward ("truth") and the plain fp32 CPU forward (the reference).

    python scripts/precision_study.py [n_windows]
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from zenker_audio_detection_b200 import synth  # noqa: E402
from oracle import numerics, thirdparty  # noqa: E402

PFX = "audio_spectrogram_transformer."


def rnd(t, fmt):
    if fmt == "bf16":
        return t.to(torch.bfloat16).to(t.dtype)
    if fmt == "fp16":
        return t.to(torch.float16).to(t.dtype)
    raise ValueError(fmt)


def split_planes(t, fmt, planes):
    out, r = [], t
    for _ in range(planes):
        p = rnd(r.float(), fmt).to(t.dtype)
        out.append(p)
        r = r - p
    return out


def make_mm(mode):
    """mode: 'f64' | 'f32' | 'bf16' | 'fp16' | 'fp16x2' (3 products) | 'bf16x3' (6 products) | 'bf16x2' (3 products)"""
    if mode == "f64":
        return lambda a, b, weight=False: a @ b
    if mode == "f32":
        return lambda a, b, weight=False: (a.float() @ b.float()).double()
    if mode in ("bf16", "fp16"):
        return lambda a, b, weight=False: rnd(a.float(), mode).double() @ rnd(b.float(), mode).double()
    scaled = mode.endswith("s")  # weights pre-scaled by a power of two so that max |w| lands in (2^13, 2^14]
    fmt, planes = mode.rstrip("s").split("x")
    planes = int(planes)

    def mm(a, b, weight=False):
        sc = 1.0
        if scaled and weight:
            sc = 2.0 ** np.floor(np.log2(16384.0 / float(b.abs().max())))
        ap, bp = split_planes(a.float().double(), fmt, planes), split_planes((b * sc).float().double(), fmt, planes)
        acc = 0
        for i in range(planes):
            for j in range(planes):
                if i + j < planes:  # drop products below the representation error
                    acc = acc + ap[i] @ bp[j]
        return acc / sc
    return mm


def forward(sd, x, mode):
    F = torch.nn.functional
    mm = make_mm(mode)
    d = torch.float64
    g = lambda k: sd[k].to(d)
    x = x.to(d)
    B = x.shape[0]
    w = g(PFX + "embeddings.patch_embeddings.projection.weight")
    cols = F.unfold(x.unsqueeze(1).transpose(2, 3), (16, 16), stride=(10, 10)).transpose(1, 2)  # (B, 1212, 256)
    pe = mm(cols, w.reshape(768, 256).t()) + g(PFX + "embeddings.patch_embeddings.projection.bias")
    x = torch.cat([g(PFX + "embeddings.cls_token").expand(B, -1, -1), g(PFX + "embeddings.distillation_token").expand(B, -1, -1), pe], 1)
    x = x + g(PFX + "embeddings.position_embeddings")
    for l in range(12):
        p = f"{PFX}encoder.layer.{l}."
        h = F.layer_norm(x, (768,), g(p + "layernorm_before.weight"), g(p + "layernorm_before.bias"), 1e-12)
        q = mm(h, g(p + "attention.attention.query.weight").t(), weight=True) + g(p + "attention.attention.query.bias")
        k = mm(h, g(p + "attention.attention.key.weight").t(), weight=True) + g(p + "attention.attention.key.bias")
        v = mm(h, g(p + "attention.attention.value.weight").t(), weight=True) + g(p + "attention.attention.value.bias")
        q, k, v = (t.view(B, -1, 12, 64).transpose(1, 2) for t in (q, k, v))
        s = mm(q, k.transpose(2, 3)) * 0.125
        a = mm(torch.softmax(s, -1), v).transpose(1, 2).reshape(B, -1, 768)
        x = x + mm(a, g(p + "attention.output.dense.weight").t(), weight=True) + g(p + "attention.output.dense.bias")
        h = F.layer_norm(x, (768,), g(p + "layernorm_after.weight"), g(p + "layernorm_after.bias"), 1e-12)
        h = F.gelu(mm(h, g(p + "intermediate.dense.weight").t(), weight=True) + g(p + "intermediate.dense.bias"))
        x = x + mm(h, g(p + "output.dense.weight").t(), weight=True) + g(p + "output.dense.bias")
    x = F.layer_norm(x, (768,), g(PFX + "layernorm.weight"), g(PFX + "layernorm.bias"), 1e-12)
    pooled = (x[:, 0] + x[:, 1]) / 2
    pooled = F.layer_norm(pooled, (768,), g("classifier.layernorm.weight"), g("classifier.layernorm.bias"), 1e-12)
    return pooled @ g("classifier.dense.weight").t() + g("classifier.dense.bias")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["f32", "bf16", "fp16", "fp16x2", "bf16x2", "bf16x3"]
    torch.set_num_threads(8)
    sd = synth.random_state_dict(11)
    wins = synth.cfg1_windows(n)
    feats = torch.from_numpy(numerics.fx_features(wins, synth.STAGE1_MEAN, synth.STAGE1_STD))
    with torch.inference_mode():
        t0 = time.time()
        truth = forward(sd, feats, "f64")
        print(f"f64 truth {time.time() - t0:.1f}s  margins", (truth[:, 1] - truth[:, 0]).numpy().round(4))
        ref32 = numerics.ast_forward(sd, feats).double()
        print(f"{'torch fp32 (reference)':>24}: max |dlogit| vs f64 {float((ref32 - truth).abs().max()):.3e}")
        for m in modes:
            t0 = time.time()
            y = forward(sd, feats, m)
            e = (y - truth).abs()
            print(f"{m:>24}: max |dlogit| vs f64 {float(e.max()):.3e}  rms {float(e.pow(2).mean().sqrt()):.3e}   ({time.time() - t0:.0f}s)")


if __name__ == "__main__":
    main()
