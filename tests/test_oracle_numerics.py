"""Pins the oracle's numerics restatement against the installed torchaudio / transformers (the packages the
reference itself calls) and against the committed golden vectors; pins the product's constant tables and the
fbank kernel's lane arithmetic (CPU emulation of the same header) without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import numerics, thirdparty
from zenker_audio_detection_b200 import synth, tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_resample_restatement_vs_torchaudio_and_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "resample.npz"))
    r = synth.recording(0.5, 48000, seed=5)
    a = numerics.resample(r, 48000, 16000)
    assert a.shape == g["out48"].shape and np.abs(a - g["out48"]).max() <= 2e-6
    assert np.abs(thirdparty.resample(r, 48000, 16000) - g["out48"]).max() == 0.0
    r2 = synth.recording(0.25, 44100, seed=6)
    b = numerics.resample(r2, 44100, 16000)
    assert b.shape == g["out441"].shape and np.abs(b - g["out441"]).max() <= 2e-6


@pytest.mark.parametrize("orig", [48000, 44100, 32000, 96000, 8000, 22050])
def test_resample_taps_bit_identical(orig):
    import torchaudio.functional.functional as TF
    import math

    gcd = math.gcd(orig, 16000)
    ref, width = TF._get_sinc_resample_kernel(orig, 16000, gcd, dtype=torch.float32)
    taps, w, o, n = tables.sinc_resample_kernel(orig, 16000)
    assert w == width and taps.shape == (n, 2 * w + o)
    assert torch.equal(taps, ref.reshape(n, -1))
    t2, w2, _, _ = numerics.sinc_resample_kernel(orig, 16000)
    assert w2 == width and np.abs(t2 - ref.reshape(n, -1).numpy()).max() <= 1e-6


def test_mel_and_window_tables_bit_identical():
    import torchaudio.compliance.kaldi as K

    ref, _ = K.get_mel_banks(128, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)
    mel = tables.mel_banks()
    assert torch.equal(mel, ref)
    assert int((mel != 0).sum()) == 504 and int((mel != 0).sum(1).max()) <= 16
    assert int(((mel != 0).sum(1) == 0).sum()) >= 1  # at least one empty filter
    for wt in ("hanning", "povey", "hamming", "rectangular"):
        assert torch.equal(tables.feature_window(wt), K._feature_window_function(wt, 400, 0.42, torch.device("cpu"), torch.float32))
    assert tables.EPSILON == K.EPSILON.item()
    assert np.abs(numerics.mel_banks(dtype=np.float64) - ref.numpy()).max() <= 5e-5  # fp32 noise of the reference table


def test_fbank_restatement_vs_torchaudio():
    w = synth.noise_16k(3.0, seed=1)
    ref = thirdparty.kaldi_fbank(w)
    f64 = numerics.fbank(w, dtype=np.float64)
    assert ref.shape == f64.shape == (298, 128)
    assert np.abs(ref - f64).max() <= 2e-3  # the reference's own fp32 noise floor
    assert np.sqrt(np.mean((ref - f64) ** 2)) <= 5e-5
    pv = thirdparty.kaldi_fbank(w, window_type="povey")
    assert np.abs(pv - numerics.fbank(w, window_type="povey", dtype=np.float64)).max() <= 2e-3
    assert numerics.fbank(w[:399]).shape == (0, 128)


def test_fx_restatement_vs_hf_and_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "fx_cfg1.npz"))
    w = synth.cfg1_windows(64)[:4]
    fx = thirdparty.hf_feature_extractor(synth.STAGE1_MEAN, synth.STAGE1_STD)
    ref = fx(list(w), sampling_rate=16000, return_tensors="np")["input_values"]
    assert np.array_equal(ref[:, :98], g["rows"]) and np.all(ref[:, 98:] == g["pad_value"])
    ours = numerics.fx_features(list(w), synth.STAGE1_MEAN, synth.STAGE1_STD, dtype=np.float64)
    assert np.all(ours[:, 98:] == pytest.approx(float(g["pad_value"]), abs=1e-6))
    d = np.abs(ours[:, :98] - g["rows"])
    assert np.sqrt(np.mean(d ** 2)) <= 1e-4 and d.max() <= 2e-2  # tone windows span > 20 nats of dynamic range


def test_ast_forward_restatement_vs_hf_two_layers():
    from transformers import ASTConfig, ASTForAudioClassification

    sd = synth.random_state_dict(3)
    cfg = ASTConfig(num_labels=2, num_hidden_layers=2)
    with torch.device("cpu"):
        m = ASTForAudioClassification(cfg)
    keep = {k: v for k, v in sd.items() if not any(f"layer.{l}." in k for l in range(2, 12))}
    m.load_state_dict(keep)
    m.eval()
    x = torch.randn(2, 1024, 128, generator=torch.Generator().manual_seed(0)) * 0.5
    with torch.inference_mode():
        ref = m(x).logits
        ours = numerics.ast_forward(sd, x, num_layers=2)
    assert (ref - ours).abs().max().item() <= 2e-5


def test_golden_ast_probs_are_softmax_of_logits(golden_dir):
    g = np.load(os.path.join(golden_dir, "ast_cfg1.npz"))
    for i in (1, 2):
        p = torch.softmax(torch.from_numpy(g[f"logits{i}"]), dim=1).numpy()
        assert np.abs(p - g[f"probs{i}"]).max() <= 1e-6
    frac = (g["probs1"][:, 1] > 0.5).mean()
    assert 0.1 <= frac <= 0.6  # the conditioned head bias splits the windows


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(ROOT, "tests", "native", "libzk_emu.so")
    src = os.path.join(ROOT, "tests", "native", "fbank_emulate.cpp")
    hdr = os.path.join(ROOT, "zenker_audio_detection_b200", "csrc", "zk_fbank_math.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off",
                               "-I" + os.path.dirname(hdr), src, "-o", so])
    return ctypes.CDLL(so)


def _emu_fbank(emu, wave):
    P = ctypes.POINTER(ctypes.c_float)
    win = tables.feature_window().numpy().copy()
    mel = tables.mel_banks().numpy().copy()
    wave = np.ascontiguousarray(wave, dtype=np.float32)
    m = 0 if len(wave) < 400 else 1 + (len(wave) - 400) // 160
    out = np.zeros((m, 128), np.float32)
    rc = emu.zk_emu_fbank(wave.ctypes.data_as(P), ctypes.c_long(len(wave)), win.ctypes.data_as(P), mel.ctypes.data_as(P),
                          ctypes.c_float(0.97), ctypes.c_float(tables.EPSILON), out.ctypes.data_as(P), ctypes.c_long(m))
    assert rc == 0
    return out


def test_fbank_kernel_lane_math_emulated_on_cpu(emu):
    """The kernel's FFT decomposition / real-FFT split / sparse mel, compiled for the host from the same header."""
    w = synth.noise_16k(5.0, seed=7)
    w[16000:20000] *= 1e-3
    w[30000:31000] = 0.0
    got = _emu_fbank(emu, w)
    ref = thirdparty.kaldi_fbank(w)
    f64 = numerics.fbank(w, dtype=np.float64)
    d = np.abs(got - ref)
    assert np.all(d <= 1e-4 * np.abs(ref) + 1e-3)
    assert np.mean(d <= 1e-4 * np.abs(ref)) >= 0.999
    assert np.sqrt(np.mean((got - f64) ** 2)) <= 1.5 * np.sqrt(np.mean((ref - f64) ** 2)) + 1e-6
    sil = 188  # frame 188 = samples [30080, 30480): fully inside the digital silence
    assert np.all(got[sil] == np.float32(np.log(tables.EPSILON)))
