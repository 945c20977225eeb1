"""Snippet-level evaluators: evaluate.run_inference against the reference's per-batch loop
(utils/analyze_ROC_PR_stage1.py:163-191) run through the installed HF packages on the CPU."""
import wave

import numpy as np
import pytest
import torch

from zenker_audio_detection_b200 import evaluate, synth
from zenker_audio_detection_b200.fx import ZenkerASTFeatureExtractor
from zenker_audio_detection_b200.model import ZenkerASTForAudioClassification

pytestmark = pytest.mark.gpu

MEAN, STD = -1.1509622, 3.5340312


def _snippets():
    """Ragged snippets the way the CV folds hold them: 0.3 - 1.5 s, several sharing a length."""
    w = synth.cfg1_windows(12, seed=77)
    lens = [16000, 4800, 16000, 24000, 9000, 16000, 4800, 12345, 16000, 24000, 401, 16000]
    out = []
    for i, n in enumerate(lens):
        x = w[i]
        out.append(np.ascontiguousarray(np.concatenate([x, x])[:n], dtype=np.float32))
    return out


def _model_dir(tmp_path, sd):
    from safetensors.torch import save_file
    from transformers import ASTConfig

    root = tmp_path / "fold1" / "best"
    root.mkdir(parents=True)
    ASTConfig(num_labels=2).save_pretrained(str(root))
    save_file({k: v.contiguous() for k, v in sd.items()}, str(root / "model.safetensors"))
    ZenkerASTFeatureExtractor(mean=MEAN, std=STD, max_length=1024).save_pretrained(str(root))
    return str(root)


def _write_wav(path, x, sr):
    pcm = np.clip(np.round(x * 32768.0), -32768, 32767).astype("<i2")
    with wave.open(str(path), "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(sr)
        f.writeframes(pcm.tobytes())
    return pcm.astype(np.float32) / 32768.0


def test_run_inference_matches_the_reference_loop(tmp_path):
    from oracle import thirdparty

    sd = synth.random_state_dict(5)
    root = _model_dir(tmp_path, sd)
    snippets = _snippets()
    got = evaluate.run_inference(root, snippets, 5)
    assert got.dtype == np.float32 and got.shape == (len(snippets),)
    # the reference loop: fx -> model -> softmax[:, 1], batch by batch (ref:176-190), HF fp32 on the CPU
    thirdparty.set_threads()
    model = thirdparty.hf_model_from_state_dict(sd)
    fx = thirdparty.hf_feature_extractor(MEAN, STD)
    want = thirdparty.forward_probs(model, fx, snippets, 4)[:, 1]
    print(f"run_inference: max |p1 - ref| = {np.abs(got - want).max():.3g}")
    # probabilities of an fp16-operand forward: 1e-2 in the logits is 2.5e-3 in p at worst
    assert np.abs(got - want).max() < 2.5e-3
    # the batch size does not change a single bit (per-window math is batch-composition independent)
    assert np.array_equal(got, evaluate.run_inference(root, snippets, 12))
    assert evaluate.run_inference(root, [], 8).shape == (0,)


def test_entries_may_be_paths_dicts_or_arrays(tmp_path):
    sd = synth.random_state_dict(6)
    fx = ZenkerASTFeatureExtractor(mean=MEAN, std=STD, max_length=1024)
    model = ZenkerASTForAudioClassification({"max_length": 1024}, sd).to("cuda")
    x = _snippets()[0]
    q = _write_wav(tmp_path / "a.wav", x, 16000)          # what the file really holds (16-bit quantised)
    x48 = synth.recording(1.0, 48000, seed=9)
    q48 = _write_wav(tmp_path / "b.wav", x48, 48000)
    entries = [q, str(tmp_path / "a.wav"), {"array": q, "sampling_rate": 16000}, {"audio": q.tolist()},
               str(tmp_path / "b.wav"), {"values": q48, "sampling_rate_hz": 48000}]
    logits = evaluate.snippet_logits(model, fx, entries, batch_size=4)
    assert logits.shape == (6, 2) and logits.dtype == np.float32
    assert np.array_equal(logits[0], logits[1]) and np.array_equal(logits[0], logits[2]) and np.array_equal(logits[0], logits[3])
    assert np.allclose(logits[4], logits[5], atol=2e-3)    # PCM16 -> GPU resampler vs float32 -> GPU resampler
    # the features are the extractor's own, snippet by snippet
    feats = evaluate.snippet_features(fx, entries[:2] + [x48[:4800]])
    one = fx([q, q, x48[:4800]], sampling_rate=16000, return_tensors="pt")["input_values"]
    assert torch.equal(feats, one)
    with pytest.raises(ValueError, match="Unsupported dict payload"):
        evaluate.to_waveform({"foo": 1})
    with pytest.raises(TypeError, match="Unsupported audio payload type"):
        evaluate.to_waveform(3.5)
