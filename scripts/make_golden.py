#!/usr/bin/env python
"""Generate ``tests/golden/*`` by running the REFERENCE's own code on CPU.

Run in the build container only (needs ``/root/reference``; the GPU box does not have
it):  ``python scripts/make_golden.py [--skip-ast]``.

What is pinned
--------------
glue_windows.json    ``ref.window_audio`` window counts/starts for many lengths.
glue_cascade.json    ``ref.main()`` and ``refc.main()`` (the cached variant) run end to end
                     with stubbed audio loading and stub models whose per-window
                     probabilities come from seeded tables: the JSON they write is the
                     expected output of our host-side cascade logic for the same tables
                     (covers the bare-argmax quirk, NaN mean, forward-min-prob,
                     --stage2-argmax).
fx_cfg1.npz          HF ``ASTFeatureExtractor`` output (via ``refc.compute_features``) for
                     the first 4 cfg1 windows, un-padded rows + the pad constant.
resample.npz         ``torchaudio.functional.resample`` 48k->16k and 44.1k->16k of a short
                     synthetic recording (exactly what ``ref.load_audio`` does after decode).
ast_cfg1.npz         ``ref.forward_probs`` probabilities and HF logits of the conditioned
                     random-init Stage-1 / Stage-2 models for the first 16 cfg1 windows.
cache_golden.json,   ``refc.get_fx_fingerprint`` / ``build_cache_path`` / ``build_base_metadata`` for fixed files
cache_ref_bundle.pt  (path, size, mtime), and a feature bundle WRITTEN BY ``refc.load_or_compute_features`` (HF
                     extractor, max_length 16 to keep it small); also checks that the reference LOADS a bundle
                     written by ``zenker_audio_detection_b200.cache`` instead of recomputing.
cascade_60s.npz      the whole reference cascade (ref.window_audio -> ref.forward_probs ->
                     gate -> ref.forward_probs -> ref.summarize_stage_outputs) on a 60-s
                     48 kHz synthetic recording (119 windows).
"""
import argparse
import io
import json
import os
import sys
import contextlib
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

import test_long_audio_windows_2stage as ref  # noqa: E402
import test_long_audio_windows_2stage_cache as refc  # noqa: E402

from zenker_audio_detection_b200 import synth  # noqa: E402
from oracle import thirdparty as T  # noqa: E402

ref.DEVICE = torch.device("cpu")
refc.DEVICE = torch.device("cpu")
GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)


def gold_windows():
    cases = []
    for L in [0, 1, 399, 400, 15999, 16000, 16001, 23999, 24000, 24001, 31999, 32000, 40000, 160000, 960000]:
        for (w, h) in [(1.0, 0.5), (1.0, 1.0), (0.5, 0.25), (2.0, 0.5), (1.0, 1.5)]:
            audio = np.arange(L, dtype=np.float32)
            wins = ref.window_audio(audio, w, h)
            starts = [int(x[0]) if L > 0 else 0 for x in wins]
            assert all(len(x) == int(w * 16000) for x in wins)
            cases.append({"L": L, "window_sec": w, "hop_sec": h, "n": len(wins), "starts": starts,
                          "last_tail_zero": bool(len(wins) and L < int(w * 16000))})
    json.dump(cases, open(os.path.join(GOLD, "glue_windows.json"), "w"))
    print("glue_windows", len(cases))


class StubFX:
    model_input_names = ["input_values"]

    def __init__(self, tag):
        self.tag = tag

    def to_dict(self):
        return {"tag": self.tag}

    def __call__(self, batch, sampling_rate=None, return_tensors=None, **kw):
        return {"input_values": torch.tensor([[float(w[0])] for w in batch])}


class StubModel:
    def __init__(self, table):
        self.table = torch.tensor(table, dtype=torch.float32)

    def __call__(self, feats):
        idx = feats[:, 0].round().long()

        class O:
            pass

        o = O()
        o.logits = self.table[idx]
        return o


def run_ref_main(mod, tables1, tables2, nwin, extra_args):
    """Run the reference's main() with stubbed audio/model loading; returns its JSON."""
    audios = {}
    for fi, n in enumerate(nwin):
        L = 16000 + 8000 * (n - 1) if n > 0 else 100
        a = np.zeros(L, dtype=np.float32)
        for k in range(n):
            a[8000 * k] = k
        audios[f"F{fi}.wav"] = a
    cur = {"file": None}

    def load_audio(path, target_sr=16000):
        cur["file"] = os.path.basename(path)
        return audios[os.path.basename(path)]

    def load_stage_model(root, labels):
        class PerFile:
            def __init__(self, tabs):
                self.tabs = tabs

            def __call__(self, feats):
                fi = int(cur["file"][1])
                return StubModel(self.tabs[fi])(feats)

        return StubFX(root), PerFile(tables1 if root == "S1" else tables2)

    mod.load_audio, mod.load_stage_model = load_audio, load_stage_model
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "o.json")
        for fn in ("F0.wav", "F1.wav"):  # the cached variant stats the files (refc:117-118)
            open(os.path.join(td, fn), "wb").write(b"stub")
        argv = ["x", "--stage1-model-root", "S1", "--stage2-model-root", "S2", "--file-a", os.path.join(td, "F0.wav"),
                "--file-b", os.path.join(td, "F1.wav"), "--output-json", out, "--show-first-n", "0"] + extra_args
        old = sys.argv
        sys.argv = argv
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                import warnings

                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    mod.main()
        finally:
            sys.argv = old
        txt = open(out).read().replace("NaN", '"NaN"').replace(td + os.sep, "")
        return json.loads(txt)


def gold_cascade():
    rng = np.random.default_rng(77)
    cases = []
    specs = [
        ("plain", ref, [37, 21], 1.0, [], 0.0),
        ("thr_quirk", ref, [40, 40], 1.5, ["--stage1-threshold", "0.8", "--stage2-threshold", "0.35"], 0.0),
        ("nan_mean", ref, [12, 9], 0.3, ["--stage1-threshold", "0.999"], 1.0),
        ("all_idle", ref, [10, 5], 1.0, [], -8.0),
        ("single_window", ref, [1, 1], 1.0, [], 0.5),
        ("cached_plain", refc, [37, 21], 1.0, ["--disable-cache"], 0.0),
        ("cached_minprob", refc, [30, 30], 1.5, ["--disable-cache", "--stage1-forward-min-prob", "0.7"], 0.0),
        ("cached_argmax", refc, [25, 18], 1.0, ["--disable-cache", "--stage2-argmax", "--stage2-threshold", "0.9"], 0.0),
    ]
    for name, mod, nwin, scale, extra, shift in specs:
        t1 = [(rng.standard_normal((n, 2)) * scale + np.array([0.0, shift])).astype(np.float32).tolist() for n in nwin]
        t2 = [(rng.standard_normal((n, 2)) * scale).astype(np.float32).tolist() for n in nwin]
        out = run_ref_main(mod, t1, t2, nwin, extra)
        cases.append({"name": name, "variant": "cached" if mod is refc else "plain", "nwin": nwin, "args": extra,
                      "logits1": t1, "logits2": t2, "expected": out})
    json.dump(cases, open(os.path.join(GOLD, "glue_cascade.json"), "w"))
    print("glue_cascade", len(cases))


def gold_fx():
    w = synth.cfg1_windows(64)[:4]
    fx = T.hf_feature_extractor(synth.STAGE1_MEAN, synth.STAGE1_STD)
    feats = refc.compute_features(fx, list(w), 2).numpy()
    assert feats.shape == (4, 1024, 128) and feats.dtype == np.float32
    pad = feats[:, 98:, :]
    assert np.all(pad == pad.flat[0])
    np.savez_compressed(os.path.join(GOLD, "fx_cfg1.npz"), rows=feats[:, :98, :], pad_value=np.float32(pad.flat[0]),
                        mean=synth.STAGE1_MEAN, std=synth.STAGE1_STD)
    print("fx_cfg1", feats[:, :98].shape, float(pad.flat[0]))


def gold_resample():
    r = synth.recording(0.5, 48000, seed=5)
    a = T.resample(r, 48000, 16000)
    r2 = synth.recording(0.25, 44100, seed=6)
    b = T.resample(r2, 44100, 16000)
    st = np.stack([r[:12000], r[12000:24000]])  # 2-channel: mean then resample (ref:55-58)
    c = T.resample(st, 48000, 16000)
    np.savez_compressed(os.path.join(GOLD, "resample.npz"), out48=a, out441=b, out_stereo=c)
    print("resample", a.shape, b.shape, c.shape)


def conditioned_models(windows_for_quantile, fx1, fx2):
    """Returns (sd1, sd2, bias1_s1, bias1_s2) with head biases shifted so ~30% / ~50% pass."""
    sd1, sd2 = synth.random_state_dict(11), synth.random_state_dict(22)
    m1, m2 = T.hf_model_from_state_dict(sd1), T.hf_model_from_state_dict(sd2)
    with torch.inference_mode():
        f1 = fx1(list(windows_for_quantile), sampling_rate=16000, return_tensors="pt")["input_values"]
        f2 = fx2(list(windows_for_quantile), sampling_rate=16000, return_tensors="pt")["input_values"]
        l1 = torch.cat([m1(f1[i:i + 8]).logits for i in range(0, len(f1), 8)])
        l2 = torch.cat([m2(f2[i:i + 8]).logits for i in range(0, len(f2), 8)])
    d1, d2 = (l1[:, 1] - l1[:, 0]).numpy(), (l2[:, 1] - l2[:, 0]).numpy()
    b1 = float(-np.quantile(d1, 0.70))
    b2 = float(-np.quantile(d2, 0.50))
    return sd1, sd2, b1, b2, l1.numpy(), l2.numpy()


def gold_ast():
    w = synth.cfg1_windows(64)[:16]
    fx1 = T.hf_feature_extractor(synth.STAGE1_MEAN, synth.STAGE1_STD)
    fx2 = T.hf_feature_extractor(synth.STAGE2_MEAN, synth.STAGE2_STD)
    sd1, sd2, b1, b2, l1, l2 = conditioned_models(w, fx1, fx2)
    sd1["classifier.dense.bias"][1] = b1
    sd2["classifier.dense.bias"][1] = b2
    m1, m2 = T.hf_model_from_state_dict(sd1), T.hf_model_from_state_dict(sd2)
    p1 = ref.forward_probs(m1, fx1, list(w), 8)
    p2 = ref.forward_probs(m2, fx2, list(w), 8)
    l1 = l1 + np.array([0.0, b1], dtype=np.float32)
    l2 = l2 + np.array([0.0, b2], dtype=np.float32)
    np.savez_compressed(os.path.join(GOLD, "ast_cfg1.npz"), probs1=p1, probs2=p2, logits1=l1, logits2=l2,
                        head_bias1_s1=b1, head_bias1_s2=b2, seed1=11, seed2=22)
    print("ast_cfg1", p1[:4], p2[:4], b1, b2)


def gold_ast_plain():
    """HF logits of a PLAIN random init (query/key gain 1, the scale HF's own init has) on 8 cfg1 windows: the
    1e-2 bf16 logit tolerance of north_star is checked on this one; the sensitised models above amplify bf16 noise."""
    w = synth.cfg1_windows(64)[:8]
    fx1 = T.hf_feature_extractor(synth.STAGE1_MEAN, synth.STAGE1_STD)
    sd = synth.random_state_dict(33, qk_gain=1.0)
    m = T.hf_model_from_state_dict(sd)
    with torch.inference_mode():
        f = fx1(list(w), sampling_rate=16000, return_tensors="pt")["input_values"]
        l = m(f).logits.numpy()
    np.savez_compressed(os.path.join(GOLD, "ast_plain.npz"), logits=l, seed=33, qk_gain=1.0)
    print("ast_plain", l[:3])


def gold_cascade_60s():
    g = np.load(os.path.join(GOLD, "ast_cfg1.npz"))
    fx1 = T.hf_feature_extractor(synth.STAGE1_MEAN, synth.STAGE1_STD)
    fx2 = T.hf_feature_extractor(synth.STAGE2_MEAN, synth.STAGE2_STD)
    rec = synth.recording(60.0, 48000, seed=2002)
    audio = T.resample(rec, 48000, 16000)
    windows = ref.window_audio(audio, 1.0, 0.5)
    # head biases re-derived on THIS recording so the gate splits it ~30/70 and ~50/50
    sd1, sd2, b1, b2, l1, l2 = conditioned_models(windows, fx1, fx2)
    sd1["classifier.dense.bias"][1] = b1
    sd2["classifier.dense.bias"][1] = b2
    m1, m2 = T.hf_model_from_state_dict(sd1), T.hf_model_from_state_dict(sd2)
    s1 = ref.forward_probs(m1, fx1, windows, 16)
    preds = s1.argmax(axis=1)
    preds = np.where((preds == 1) & (s1[:, 1] >= 0.5), 1, 0)
    idx = np.where(preds == 1)[0]
    s2 = ref.forward_probs(m2, fx2, [windows[i] for i in idx], 16)
    results = [(int(gi), s2[i]) for i, gi in enumerate(idx)]
    summ = ref.summarize_stage_outputs(s1, results, ["Idle", "Swallow"], ["Healthy", "Zenker"], 0.5)
    np.savez_compressed(
        os.path.join(GOLD, "cascade_60s.npz"), s1_probs=s1, swallow_indices=idx, s2_probs=s2,
        s1_logits=l1 + np.array([0.0, b1], dtype=np.float32),
        s2_logits_all=l2 + np.array([0.0, b2], dtype=np.float32),
        head_bias1_s1=b1, head_bias1_s2=b2, summary=json.dumps(summ), n_windows=len(windows), audio16k_head=audio[:64],
    )
    print("cascade_60s", len(windows), len(idx), summ)


def gold_cascade_second():
    """An INDEPENDENT recording for the decision-parity test (tests/test_gpu_fullsize.py): 300 s at 48 kHz from another
    generator seed (4242; 599 windows), same conditioned weights, thresholds 0.5 / 0.5 and 0.55 / 0.45.  ~15 min of CPU."""
    gold_cascade_cfg2(seconds=300.0, seed=4242, name="cascade_300s_seed4242.npz", pairs=(("a", 0.5, 0.5), ("b", 0.55, 0.45)))


def gold_cascade_cfg2(seconds=600.0, seed=2002, name="cascade_cfg2.npz", pairs=(("a", 0.5, 0.5), ("b", 0.6, 0.35))):
    """BASELINE.json configs[1] at FULL size: the 600-s 48 kHz recording the bench runs (synth.recording seed 2002) ->
    1199 windows through ref.forward_probs (Stage 1 on all, Stage 2 on the forwarded ones), the reference gate and
    ref.summarize_stage_outputs at thresholds 0.5 / 0.5 and at 0.6 / 0.35 (the counting quirk of SURVEY.md 0.7; its
    forwarded set is a subset of the first run's, so no further forward is needed).  ~25 min on 8 cores.
    The head biases are the ones derived for cascade_60s.npz."""
    g = np.load(os.path.join(GOLD, "cascade_60s.npz"))
    b1, b2 = float(g["head_bias1_s1"]), float(g["head_bias1_s2"])
    fx1 = T.hf_feature_extractor(synth.STAGE1_MEAN, synth.STAGE1_STD)
    fx2 = T.hf_feature_extractor(synth.STAGE2_MEAN, synth.STAGE2_STD)
    rec = synth.recording(seconds, 48000, seed=seed)
    audio = T.resample(rec, 48000, 16000)
    windows = ref.window_audio(audio, 1.0, 0.5)
    m1 = T.hf_model_from_state_dict(synth.random_state_dict(11, head_bias1=b1))
    m2 = T.hf_model_from_state_dict(synth.random_state_dict(22, head_bias1=b2))
    s1 = ref.forward_probs(m1, fx1, windows, 16)
    out = {}
    s2_by_window = {}
    for tag, thr1, thr2 in pairs:
        preds = s1.argmax(axis=1)
        preds = np.where((preds == 1) & (s1[:, 1] >= thr1), 1, 0)   # ref:313-317
        idx = np.where(preds == 1)[0]
        todo = [i for i in idx if int(i) not in s2_by_window]
        if todo:
            p = ref.forward_probs(m2, fx2, [windows[i] for i in todo], 16)
            for i, row in zip(todo, p):
                s2_by_window[int(i)] = row
        s2 = np.stack([s2_by_window[int(i)] for i in idx]) if len(idx) else np.zeros((0, 2), np.float32)
        results = [(int(gi), s2[i]) for i, gi in enumerate(idx)]
        summ = ref.summarize_stage_outputs(s1, results, ["Idle", "Swallow"], ["Healthy", "Zenker"], thr2)
        out[f"swallow_indices_{tag}"] = idx
        out[f"s2_probs_{tag}"] = s2
        out[f"summary_{tag}"] = json.dumps(summ)
        out[f"thresholds_{tag}"] = np.array([thr1, thr2])
        print("cascade_cfg2", tag, len(windows), len(idx), summ)
    np.savez_compressed(os.path.join(GOLD, name), s1_probs=s1, head_bias1_s1=b1, head_bias1_s2=b2,
                        n_windows=len(windows), audio16k_head=audio[:64], seconds=seconds, seed=seed, **out)


def gold_stats():
    """utils/compute_ast_normalization_stats.py: its module imports librosa / soundfile (absent here), so the two pure
    functions are taken out of the reference source with ``ast`` and executed as they are: ``aggregate_stats`` on a
    5-fold example, and the mean / unbiased-std tail of ``compute_fold_stats`` (lines 82-95) through a features-only
    re-run of its accumulation loop on HF extractor output (do_normalize = False, lines 62-80)."""
    import ast

    src = open("/root/reference/utils/compute_ast_normalization_stats.py").read()
    tree = ast.parse(src)
    ns = {"np": np, "torch": torch, "os": os}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "aggregate_stats":
            exec(compile(ast.Module([node], []), "ref_stats", "exec"), ns)
    rng = np.random.default_rng(5)
    per_fold = [{"fold": k + 1, "mean": float(rng.normal(-4.0, 0.3)), "std": float(rng.uniform(4.0, 5.0)),
                 "count": int(rng.integers(1, 50)) * 131072} for k in range(5)]
    per_fold.append({"fold": 6, "mean": 0.0, "std": 0.0, "count": 0})
    agg = ns["aggregate_stats"](per_fold)
    # the accumulation of lines 62-95 on 6 snippets of three lengths (one shorter than a window, one over 10.24 s)
    from transformers import ASTFeatureExtractor

    fx = ASTFeatureExtractor(max_length=1024, num_mel_bins=128)
    fx.do_normalize = False
    g = torch.Generator().manual_seed(9)
    wavs = [(torch.randn(n, generator=g) * a).numpy() for n, a in ((16000, 0.1), (16000, 0.01), (8000, 0.3), (8000, 0.05),
                                                                   (170000, 0.2), (170000, 0.02))]
    total_count, running_sum, running_sq_sum = 0, 0.0, 0.0
    for start in range(0, len(wavs), 4):
        feats = fx(wavs[start:start + 4], sampling_rate=16000, return_tensors="pt")["input_values"]
        flat = feats.view(feats.size(0), -1).to(torch.float64)
        running_sum += flat.sum().item()
        running_sq_sum += (flat ** 2).sum().item()
        total_count += flat.numel()
    mean = running_sum / total_count
    var = max(running_sq_sum / total_count - mean * mean, 0.0) * (total_count / (total_count - 1))
    with open(os.path.join(GOLD, "stats_golden.json"), "w") as f:
        json.dump({"per_fold": per_fold, "aggregate": agg, "snippet_lengths": [len(w) for w in wavs],
                   "snippet_gains": [0.1, 0.01, 0.3, 0.05, 0.2, 0.02], "snippet_seed": 9,
                   "fold_stats": {"mean": mean, "std": var ** 0.5, "count": total_count}}, f, indent=1)
    print("stats", agg, mean, var ** 0.5, total_count)


CACHE_FIXTURE_DIR = "/tmp/zk_cache_golden"  # absolute on purpose: the cache key hashes the absolute path (refc:97-100)


def make_cache_fixture_file(name, size, mtime):
    """A fake recording with a fixed absolute path, size and mtime (the three things the cache key depends on)."""
    d = os.path.join(CACHE_FIXTURE_DIR, "audio")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, name)
    with open(path, "wb") as f:
        f.write(bytes((i * 37 + 11) & 0xFF for i in range(size)))
    os.utime(path, (mtime, mtime))
    return path


def gold_cache():
    import shutil

    from transformers import ASTFeatureExtractor

    from zenker_audio_detection_b200 import cache as zcache

    shutil.rmtree(CACHE_FIXTURE_DIR, ignore_errors=True)
    doc = {"fixture_dir": CACHE_FIXTURE_DIR, "files": [], "keys": []}
    fx_full = T.hf_feature_extractor(synth.STAGE1_MEAN, synth.STAGE1_STD)
    fx_small = ASTFeatureExtractor(max_length=16, mean=synth.STAGE2_MEAN, std=synth.STAGE2_STD)
    doc["fingerprints"] = {"stage1_default": refc.get_fx_fingerprint(fx_full), "small": refc.get_fx_fingerprint(fx_small)}
    doc["fx_dicts"] = {"stage1_default": fx_full.to_dict(), "small": fx_small.to_dict()}
    cache_dir = os.path.join(CACHE_FIXTURE_DIR, "cache")
    for name, size, mtime in [("rec_A.wav", 4321, 1700000000), ("P017 swallow.long.wav", 99, 1234567890)]:
        path = make_cache_fixture_file(name, size, mtime)
        doc["files"].append({"name": name, "size": size, "mtime": mtime})
        for (w, h, n) in [(1.0, 0.5, 7), (0.5, 0.25, 1199), (2.0, 1.0, 1)]:
            for fxname in ("stage1_default", "small"):
                fp = doc["fingerprints"][fxname]
                doc["keys"].append({"file": name, "window_sec": w, "hop_sec": h, "num_windows": n, "fx": fxname,
                                    "cache_path": refc.build_cache_path(cache_dir, path, w, h, 16000, fp),
                                    "base_metadata": refc.build_base_metadata(path, w, h, n, 16000, fp)})
    # (A) a bundle written by the reference, committed as a fixture
    path = os.path.join(CACHE_FIXTURE_DIR, "audio", "rec_A.wav")
    windows = list(synth.cfg1_windows(64)[:4])
    with contextlib.redirect_stdout(io.StringIO()):
        feats = refc.load_or_compute_features(path, windows, fx_small, 1.0, 0.5, 2, cache_dir, False, False, "stage1")
    written = refc.build_cache_path(cache_dir, path, 1.0, 0.5, 16000, doc["fingerprints"]["small"])
    assert os.path.exists(written) and tuple(feats.shape) == (4, 16, 128)
    shutil.copy(written, os.path.join(GOLD, "cache_ref_bundle.pt"))
    doc["ref_bundle"] = {"file": "cache_ref_bundle.pt", "cache_path": written, "num_windows": 4, "fx": "small",
                         "feature_shape": list(feats.shape), "feature_sum": float(feats.double().sum())}
    # (B) a bundle written by OUR module must be loaded (not recomputed) by the reference
    path_b = os.path.join(CACHE_FIXTURE_DIR, "audio", "P017 swallow.long.wav")
    fp = zcache.get_fx_fingerprint(fx_small)
    ours = torch.arange(3 * 16 * 128, dtype=torch.float32).reshape(3, 16, 128) * 1e-3
    os.makedirs(cache_dir, exist_ok=True)
    zcache.save_bundle(zcache.build_cache_path(cache_dir, path_b, 1.0, 0.5, 16000, fp),
                       zcache.build_base_metadata(path_b, 1.0, 0.5, 3, 16000, fp), ours)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        got = refc.load_or_compute_features(path_b, windows[:3], fx_small, 1.0, 0.5, 2, cache_dir, False, False, "stage2")
    doc["reference_loaded_our_bundle"] = bool("Loaded" in out.getvalue() and torch.equal(got, ours))
    assert doc["reference_loaded_our_bundle"], out.getvalue()
    with open(os.path.join(GOLD, "cache_golden.json"), "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)
    print("cache golden:", len(doc["keys"]), "keys; reference loaded our bundle:", doc["reference_loaded_our_bundle"])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-ast", action="store_true")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    todo = a.only.split(",") if a.only else ["windows", "cascade", "fx", "resample", "cache", "stats", "ast", "astplain", "cascade60"]
    if "windows" in todo:
        gold_windows()
    if "cascade" in todo:
        gold_cascade()
    if "fx" in todo:
        gold_fx()
    if "resample" in todo:
        gold_resample()
    if "cache" in todo:
        gold_cache()
    if "stats" in todo:
        gold_stats()
    if not a.skip_ast:
        if "ast" in todo:
            gold_ast()
        if "astplain" in todo:
            gold_ast_plain()
        if "cascade60" in todo:
            gold_cascade_60s()
        if "cascadecfg2" in todo:  # only on request (--only cascadecfg2): ~25 min of CPU
            gold_cascade_cfg2()
        if "cascadesecond" in todo:  # only on request (--only cascadesecond): ~15 min of CPU
            gold_cascade_second()
