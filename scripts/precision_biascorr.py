"""Would a bias correction of the rounded weights narrow the re-check band?  (design study)

The fast path's margin error is mostly the 11-bit weights (scripts/precision_sites.py), and nearly the same offset for
every window -- i.e. the part  mean(x) . (W - fp16(W))  of the product, which a per-output-feature constant can absorb
(the "bias correction" of post-training quantisation).  Here: mean operand vectors of every GEMM are taken from the
float64 forward of CALIBRATION windows, the corrected biases are applied to the all-fp16 forward of OTHER windows, and the
margin error against float64 is compared with and without.

    python scripts/precision_biascorr.py [n_eval] [n_calib]
"""
import sys
import time

import torch

sys.path.insert(0, ".")
from zenker_audio_detection_b200 import synth  # noqa: E402
from oracle import numerics  # noqa: E402

PFX = "audio_spectrogram_transformer."


def forward(sd, x, rounded, means=None, corr=None):
    F = torch.nn.functional
    r = (lambda t: t.float().half().double()) if rounded else (lambda t: t)
    d = torch.float64
    g = lambda k: sd[k].to(d)
    x = x.to(d)
    B = x.shape[0]

    def lin(inp, wkey, bkey, tag):
        w = g(wkey).t()
        if means is not None:
            means[tag] = means.get(tag, 0) + inp.reshape(-1, inp.shape[-1]).mean(0) / 1.0
        y = r(inp) @ r(w) + g(bkey)
        if corr is not None:
            y = y + corr[tag] @ (w - r(w))
        return y

    w = g(PFX + "embeddings.patch_embeddings.projection.weight")
    cols = F.unfold(x.unsqueeze(1).transpose(2, 3), (16, 16), stride=(10, 10)).transpose(1, 2)
    pe = r(cols) @ r(w.reshape(768, 256).t()) + g(PFX + "embeddings.patch_embeddings.projection.bias")
    x = torch.cat([g(PFX + "embeddings.cls_token").expand(B, -1, -1), g(PFX + "embeddings.distillation_token").expand(B, -1, -1), pe], 1)
    x = x + g(PFX + "embeddings.position_embeddings")
    for l in range(12):
        p = f"{PFX}encoder.layer.{l}."
        h = F.layer_norm(x, (768,), g(p + "layernorm_before.weight"), g(p + "layernorm_before.bias"), 1e-12)
        q, k, v = (lin(h, p + f"attention.attention.{n}.weight", p + f"attention.attention.{n}.bias", f"{l}.{n}")
                   for n in ("query", "key", "value"))
        q, k, v = (r(t).view(B, -1, 12, 64).transpose(1, 2) for t in (q, k, v))
        s = (q @ k.transpose(2, 3)) * 0.125
        e = torch.exp(s - s.amax(-1, keepdim=True))
        a = ((r(e) @ v) / e.sum(-1, keepdim=True)).transpose(1, 2).reshape(B, -1, 768)
        x = x + lin(a, p + "attention.output.dense.weight", p + "attention.output.dense.bias", f"{l}.out")
        h = F.layer_norm(x, (768,), g(p + "layernorm_after.weight"), g(p + "layernorm_after.bias"), 1e-12)
        h = F.gelu(lin(h, p + "intermediate.dense.weight", p + "intermediate.dense.bias", f"{l}.fc1"))
        x = x + lin(h, p + "output.dense.weight", p + "output.dense.bias", f"{l}.fc2")
    x = F.layer_norm(x, (768,), g(PFX + "layernorm.weight"), g(PFX + "layernorm.bias"), 1e-12)
    pooled = (x[:, 0] + x[:, 1]) / 2
    pooled = F.layer_norm(pooled, (768,), g("classifier.layernorm.weight"), g("classifier.layernorm.bias"), 1e-12)
    y = pooled @ g("classifier.dense.weight").t() + g("classifier.dense.bias")
    return y[:, 1] - y[:, 0]


def main():
    n_eval = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    n_cal = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    torch.set_num_threads(8)
    sd = synth.random_state_dict(11)
    feats = torch.from_numpy(numerics.fx_features(synth.cfg1_windows(n_eval + n_cal), synth.STAGE1_MEAN, synth.STAGE1_STD))
    ev, cal = feats[:n_eval], feats[n_eval:]
    with torch.inference_mode():
        t0 = time.time()
        means = {}
        forward(sd, cal, False, means=means)   # float64 forward of the calibration windows: mean operand of every GEMM
        truth = forward(sd, ev, False)
        print(f"truth + calibration in {time.time() - t0:.0f}s", flush=True)
        for name, corr in (("fp16 operands", None), ("fp16 operands + bias correction", means)):
            e = (forward(sd, ev, True, corr=corr) - truth).abs()
            print(f"{name:>34}: max {float(e.max()):.2e}  rms {float(e.pow(2).mean().sqrt()):.2e}", flush=True)


if __name__ == "__main__":
    main()
