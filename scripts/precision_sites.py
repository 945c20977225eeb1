"""Which operand rounding carries the fast path's margin error?  (design study for DESIGN.md section 4b)

The AST forward in float64 with ONE class of matmul operands at a time rounded to fp16 (everything else exact), then all
of them, on the conditioned test weights; the figure is max / rms |(l1 - l0) - truth| over the windows.  Sites:
activations ``h1`` (LayerNorm -> QKV), ``q``, ``k``, ``p`` (softmax probabilities), ``v``, ``a`` (attention output ->
out-projection), ``h2`` (LayerNorm -> fc1), ``g`` (GELU -> fc2), ``patch``; weights ``Wqkv``, ``Wo``, ``W1``, ``W2``.

    python scripts/precision_sites.py [n_windows] [fmt]
"""
import sys
import time

import torch

sys.path.insert(0, ".")
from zenker_audio_detection_b200 import synth  # noqa: E402
from oracle import numerics  # noqa: E402

PFX = "audio_spectrogram_transformer."
SITES = ["patch", "h1", "Wqkv", "q", "k", "p", "v", "a", "Wo", "h2", "W1", "g", "W2"]


def forward(sd, x, on, fmt):
    F = torch.nn.functional
    dt = torch.float16 if fmt == "fp16" else torch.bfloat16
    r = lambda t, site: t.float().to(dt).double() if site in on else t
    d = torch.float64
    g = lambda k: sd[k].to(d)
    x = x.to(d)
    B = x.shape[0]
    w = g(PFX + "embeddings.patch_embeddings.projection.weight")
    cols = F.unfold(x.unsqueeze(1).transpose(2, 3), (16, 16), stride=(10, 10)).transpose(1, 2)
    pe = r(cols, "patch") @ r(w.reshape(768, 256).t(), "patch") + g(PFX + "embeddings.patch_embeddings.projection.bias")
    x = torch.cat([g(PFX + "embeddings.cls_token").expand(B, -1, -1), g(PFX + "embeddings.distillation_token").expand(B, -1, -1), pe], 1)
    x = x + g(PFX + "embeddings.position_embeddings")
    for l in range(12):
        p = f"{PFX}encoder.layer.{l}."
        h = r(F.layer_norm(x, (768,), g(p + "layernorm_before.weight"), g(p + "layernorm_before.bias"), 1e-12), "h1")
        q, k, v = (h @ r(g(p + f"attention.attention.{n}.weight").t(), "Wqkv") + g(p + f"attention.attention.{n}.bias")
                   for n in ("query", "key", "value"))
        q, k, v = (t.view(B, -1, 12, 64).transpose(1, 2) for t in (r(q, "q"), r(k, "k"), r(v, "v")))
        s = (q @ k.transpose(2, 3)) * 0.125
        # the kernel rounds the UN-normalised exponentials (relative rounding, like this) and divides by the fp32 row sum
        e = torch.exp(s - s.amax(-1, keepdim=True))
        a = (r(e, "p") @ v) / e.sum(-1, keepdim=True)
        a = r(a.transpose(1, 2).reshape(B, -1, 768), "a")
        x = x + a @ r(g(p + "attention.output.dense.weight").t(), "Wo") + g(p + "attention.output.dense.bias")
        h = r(F.layer_norm(x, (768,), g(p + "layernorm_after.weight"), g(p + "layernorm_after.bias"), 1e-12), "h2")
        h = r(F.gelu(h @ r(g(p + "intermediate.dense.weight").t(), "W1") + g(p + "intermediate.dense.bias")), "g")
        x = x + h @ r(g(p + "output.dense.weight").t(), "W2") + g(p + "output.dense.bias")
    x = F.layer_norm(x, (768,), g(PFX + "layernorm.weight"), g(PFX + "layernorm.bias"), 1e-12)
    pooled = (x[:, 0] + x[:, 1]) / 2
    pooled = F.layer_norm(pooled, (768,), g("classifier.layernorm.weight"), g("classifier.layernorm.bias"), 1e-12)
    y = pooled @ g("classifier.dense.weight").t() + g("classifier.dense.bias")
    return y[:, 1] - y[:, 0]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    fmt = sys.argv[2] if len(sys.argv) > 2 else "fp16"
    torch.set_num_threads(8)
    sd = synth.random_state_dict(11)
    feats = torch.from_numpy(numerics.fx_features(synth.cfg1_windows(n), synth.STAGE1_MEAN, synth.STAGE1_STD))
    with torch.inference_mode():
        t0 = time.time()
        truth = forward(sd, feats, (), fmt)
        print(f"{n} windows, {fmt}; f64 truth in {time.time() - t0:.0f}s; margins {truth.numpy().round(3)}", flush=True)
        rows = [(s, (s,)) for s in SITES] + [("all", tuple(SITES)), ("activations", ("patch", "h1", "q", "k", "p", "v", "a", "h2", "g")),
                                             ("weights", ("Wqkv", "Wo", "W1", "W2")), ("q+k", ("q", "k"))]
        for name, on in rows:
            e = (forward(sd, feats, on, fmt) - truth).abs()
            print(f"{name:>12}: max {float(e.max()):.2e}  rms {float(e.pow(2).mean().sqrt()):.2e}", flush=True)


if __name__ == "__main__":
    main()
