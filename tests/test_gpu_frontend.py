"""Resampler and Kaldi fbank (continuous + ASTFeatureExtractor contract) vs the installed torchaudio /
transformers executed on the CPU, and vs the committed golden fixtures."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

FLOOR = float(np.log(np.finfo(np.float32).eps))


EPS32 = float(np.finfo(np.float32).eps)


def fbank_tol(ref):
    """Element-wise tolerance between two independent fp32 fbank implementations:
    1e-4 relative + 1e-3 absolute, plus the fp32 noise floor of a bin that sits far below its frame's peak: an
    fp32 FFT carries ~eps * peak amplitude of error, i.e. d(log P) ~ 2 eps sqrt(P_peak / P) = 2 eps exp((max-ref)/2);
    the reference (torchaudio in fp32) is itself only this accurate against a float64 evaluation."""
    peak = ref.max(axis=-1, keepdims=True)
    return 1e-4 * np.abs(ref) + 1e-3 + EPS32 * np.exp(np.minimum((peak - ref) / 2.0, 40.0))


def fbank_gate(got, ref, ref64=None):
    d = np.abs(got - ref)
    assert np.all(d <= fbank_tol(ref)), float((d - fbank_tol(ref)).max())
    frac = float(np.mean(d <= 1e-4 * np.abs(ref)))
    assert frac >= 0.999, frac
    if ref64 is not None and got.shape[0] >= 50:  # our error against float64 is no worse than the reference's own
        rms_ours, rms_ref = np.sqrt(np.mean((got - ref64) ** 2)), np.sqrt(np.mean((ref - ref64) ** 2))
        assert rms_ours <= 1.5 * rms_ref + 1e-6, (rms_ours, rms_ref)
        assert np.abs(got - ref64).max() <= 4.0 * np.abs(ref - ref64).max() + 1e-4


@pytest.mark.parametrize("seconds", [0.025, 0.03, 1.0, 7.3, 60.0])
def test_fbank_continuous(seconds):
    from oracle import numerics, thirdparty
    from zenker_audio_detection_b200 import ops, synth

    wave = synth.noise_16k(seconds, seed=int(seconds * 1000) + 1)
    if seconds > 5:
        wave[16000:32000] *= 0.001  # quiet stretch
        wave[40000:41000] = 0.0     # digital silence -> log floor everywhere (frame 251 = [40160, 40560))
    plan = ops.FbankPlan()
    got = plan.fbank(torch.from_numpy(wave).cuda()).cpu().numpy()
    ref = thirdparty.kaldi_fbank(wave)
    assert got.shape == ref.shape
    fbank_gate(got, ref, numerics.fbank(wave, dtype=np.float64) if seconds <= 8 else None)
    # empty mel filters always sit at the floor
    empty = np.where(np.all(ref == ref[0:1, :], axis=0) & (np.abs(ref[0] - FLOOR) < 1e-5))[0]
    assert len(empty) >= 1
    for c in empty:
        assert np.all(got[:, c] == np.float32(FLOOR))
    if seconds > 5:
        assert np.all(got[251] == np.float32(FLOOR)) and np.all(ref[251] == np.float32(FLOOR))


def test_fbank_too_short_is_empty():
    from zenker_audio_detection_b200 import ops

    plan = ops.FbankPlan()
    assert plan.fbank(torch.zeros(399, device="cuda")).shape == (0, 128)


def test_fbank_window_equals_strided_rows():
    """fbank(window k) == continuous fbank rows [50k, 50k+98) bit-exactly."""
    from zenker_audio_detection_b200 import ops, synth

    wave = torch.from_numpy(synth.recording(12.0, 16000, seed=9)).cuda()
    plan = ops.FbankPlan()
    whole = plan.fbank(wave)
    wins = torch.stack([wave[8000 * k: 8000 * k + 16000] for k in range(23)])
    feats = plan.fx_contract(wins, 0.0, 0.5, 1024, do_normalize=False)
    for k in range(23):
        assert torch.equal(feats[k, :98], whole[50 * k: 50 * k + 98])
        assert torch.all(feats[k, 98:] == 0)


def test_fx_contract_vs_hf_and_golden(golden_dir):
    from oracle import thirdparty
    from zenker_audio_detection_b200 import ops, synth

    w = synth.cfg1_windows(64)
    plan = ops.FbankPlan()
    got = plan.fx_contract(torch.from_numpy(w).cuda(), synth.STAGE1_MEAN, synth.STAGE1_STD, 1024).cpu().numpy()
    fx = thirdparty.hf_feature_extractor(synth.STAGE1_MEAN, synth.STAGE1_STD)
    ref = fx(list(w), sampling_rate=16000, return_tensors="np")["input_values"]
    assert got.shape == ref.shape == (64, 1024, 128) and got.dtype == np.float32
    assert np.array_equal(got[:, 98:], ref[:, 98:])  # pad rows hold exactly (0-mean)/(2*std)
    s2 = np.float32(2 * synth.STAGE1_STD)
    raw_ref = ref[:, :98] * s2 + np.float32(synth.STAGE1_MEAN)  # back to the un-normalised log-mel domain
    d = np.abs(got[:, :98] - ref[:, :98]) * s2
    assert np.all(d <= fbank_tol(raw_ref) + 1e-5)
    assert np.mean(d <= 1e-4 * np.abs(raw_ref) + 1e-5) >= 0.999
    gold = np.load(os.path.join(golden_dir, "fx_cfg1.npz"))
    assert np.array_equal(ref[:4, :98], gold["rows"])  # the installed HF stack still produces the committed fixture
    assert np.all(got[:4, 98:] == gold["pad_value"])


@pytest.mark.parametrize("n", [400, 16000, 16037, 5000])
def test_fx_contract_ragged_and_truncate(n):
    from oracle import thirdparty
    from zenker_audio_detection_b200 import ops, synth

    w = synth.noise_16k(n / 16000.0 * 3, seed=n)[: 3 * n].reshape(3, n)
    plan = ops.FbankPlan()
    for max_length in (1024, 20):
        got = plan.fx_contract(torch.from_numpy(w.copy()).cuda(), -4.27, 4.57, max_length).cpu().numpy()
        fx = thirdparty.hf_feature_extractor(-4.27, 4.57, max_length=max_length)
        ref = fx(list(w), sampling_rate=16000, return_tensors="np")["input_values"]
        assert got.shape == ref.shape
        assert np.all(np.abs(got - ref) <= 1e-4 * np.abs(ref) + 2e-4)


@pytest.mark.parametrize("sr,seconds,channels", [(48000, 2.0, 1), (48000, 0.5, 2), (44100, 0.7, 1), (32000, 1.0, 1),
                                                 (96000, 0.3, 1), (16000, 0.5, 2), (8000, 0.5, 1)])
def test_resample_vs_torchaudio(sr, seconds, channels):
    from oracle import thirdparty
    from zenker_audio_detection_b200 import ops, synth

    r = synth.recording(seconds * channels, sr, seed=sr // 100 + channels)
    n = len(r) // channels
    x = r[: n * channels].reshape(channels, n)
    ref = thirdparty.resample(x, sr, 16000)
    got = ops.resample(torch.from_numpy(x).cuda(), sr, 16000).cpu().numpy()
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 2e-6 + 1e-5 * np.abs(ref).max()


def test_resample_golden(golden_dir):
    from zenker_audio_detection_b200 import ops, synth

    g = np.load(os.path.join(golden_dir, "resample.npz"))
    r = synth.recording(0.5, 48000, seed=5)
    got = ops.resample(torch.from_numpy(r).cuda(), 48000, 16000).cpu().numpy()
    assert np.abs(got - g["out48"]).max() <= 2e-6
    r2 = synth.recording(0.25, 44100, seed=6)
    got2 = ops.resample(torch.from_numpy(r2).cuda(), 44100, 16000).cpu().numpy()
    assert np.abs(got2 - g["out441"]).max() <= 2e-6
    st = np.stack([r[:12000], r[12000:24000]])
    got3 = ops.resample(torch.from_numpy(st).cuda(), 48000, 16000).cpu().numpy()
    assert np.abs(got3 - g["out_stereo"].reshape(-1)).max() <= 2e-6


def test_resample_pcm16():
    from oracle import thirdparty
    from zenker_audio_detection_b200 import ops, synth

    r = synth.recording(1.0, 48000, seed=12)
    pcm = np.round(r * 32767).astype(np.int16).reshape(-1, 2)  # (n, 2) interleaved stereo
    ref = thirdparty.resample((pcm.astype(np.float32) / 32768.0).T.copy(), 48000, 16000)
    got = ops.resample(torch.from_numpy(pcm).cuda(), 48000, 16000).cpu().numpy()
    assert np.abs(got - ref).max() <= 2e-6


def test_stats_epilogue_and_cli_match_the_reference_accumulation(golden_dir, tmp_path):
    """zk_fx_stats_f32 (sums inside the feature kernel, nothing written) and `python -m ...stats` on WAV files against the
    reference's accumulation loop run on the HF extractor (golden: utils/compute_ast_normalization_stats.py:62-95)."""
    import json
    import os

    from zenker_audio_detection_b200 import ops, stats, wavio

    g = json.load(open(os.path.join(golden_dir, "stats_golden.json")))
    gen = torch.Generator().manual_seed(int(g["snippet_seed"]))
    wavs = [(torch.randn(n, generator=gen) * a) for n, a in zip(g["snippet_lengths"], g["snippet_gains"])]
    ref = g["fold_stats"]
    plan = ops.FbankPlan()
    st = ops.FeatureStats()
    feats = None
    for n in sorted(set(g["snippet_lengths"])):
        group = torch.stack([w for w in wavs if w.numel() == n]).cuda()
        out = st.update_from_waveforms(plan, group, 1024, return_features=(n == 16000))
        if out is not None:
            feats = out
    got = st.result()
    assert got["count"] == ref["count"]
    assert abs(got["mean"] - ref["mean"]) <= 2e-6 * abs(ref["mean"]) + 1e-7
    assert abs(got["std"] - ref["std"]) <= 2e-6 * ref["std"]
    # the features it can also return are the contract kernel's (un-normalised, zero padded)
    want = plan.fx_contract(torch.stack(wavs[:2]).cuda(), 0.0, 0.5, 1024, do_normalize=False)
    assert torch.equal(feats, want) and float(feats[:, 98:].abs().max()) == 0.0
    # the command-line drop-in on float32 WAV files listed in train_x_fold1.npy
    import struct

    paths = []
    for i, w in enumerate(wavs):
        p = tmp_path / f"s{i}.wav"
        b = w.numpy().astype("<f4")
        with open(p, "wb") as f:
            f.write(b"RIFF" + struct.pack("<I", 36 + b.nbytes) + b"WAVE")
            f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 3, 1, 16000, 64000, 4, 32))
            f.write(b"data" + struct.pack("<I", b.nbytes) + b.tobytes())
        paths.append(str(p))
    np.save(tmp_path / "train_x_fold1.npy", np.array(paths))
    np.save(tmp_path / "train_x_fold2.npy", np.array(paths[:2]))
    stats.main(["--data-dir", str(tmp_path), "--output-dir", str(tmp_path / "out"), "--folds", "2", "--batch-size", "4"])
    per_fold = json.load(open(tmp_path / "out" / "stats_per_fold.json"))
    assert [d["fold"] for d in per_fold] == [1, 2] and per_fold[0]["count"] == ref["count"]
    assert abs(per_fold[0]["mean"] - ref["mean"]) <= 2e-6 * abs(ref["mean"]) + 1e-7
    assert abs(per_fold[0]["std"] - ref["std"]) <= 2e-6 * ref["std"]
    agg = json.load(open(tmp_path / "out" / "stats_aggregate.json"))
    assert agg == stats.aggregate_stats(per_fold)
    assert os.path.exists(tmp_path / "out" / "stats_all.npz")
