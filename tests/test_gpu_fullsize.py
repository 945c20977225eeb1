"""BASELINE.json's full-size configurations checked through size-independent properties (the CPU oracle needs ~16 min
for one cfg2 recording, so at these sizes the checks are invariances of the path itself plus spot checks against the
oracle on sampled pieces):

cfg2  10-min 48 kHz recording, 1199 windows: window geometry (ref:62-75), the gate/compaction recomputed by the oracle
      from our probabilities (bit-exact), independence of the batch composition (bit-exact), locality (a 60-s prefix of
      the recording reproduces the first windows bit-exactly), Stage-2 probabilities of the compacted windows equal to a
      direct Stage-2 forward of those windows.
cfg3  1 h of 16 kHz audio, 359 998 frames: frame independence (a shifted view of the signal gives the same rows
      bit-exactly), per-window contract output == strided rows of the continuous fbank, sampled frames vs torchaudio.
cfg4  a batch of recordings: results do not depend on the order or the rank a recording is processed on.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pipeline(batch_size=128, head_bias1=0.0, **kw):
    from zenker_audio_detection_b200 import synth
    from zenker_audio_detection_b200.fx import ZenkerASTFeatureExtractor
    from zenker_audio_detection_b200.model import ZenkerASTForAudioClassification
    from zenker_audio_detection_b200.pipeline import TwoStagePipeline

    fx1 = ZenkerASTFeatureExtractor(mean=synth.STAGE1_MEAN, std=synth.STAGE1_STD)
    fx2 = ZenkerASTFeatureExtractor(mean=synth.STAGE2_MEAN, std=synth.STAGE2_STD)
    m1 = ZenkerASTForAudioClassification({"max_length": 1024}, synth.random_state_dict(11, head_bias1=head_bias1))
    m2 = ZenkerASTForAudioClassification({"max_length": 1024}, synth.random_state_dict(22))
    return TwoStagePipeline(m1, fx1, m2, fx2, batch_size=batch_size, **kw)


@pytest.fixture(scope="module")
def cfg2():
    from zenker_audio_detection_b200 import synth

    rec = synth.recording(600.0, 48000, seed=2002)
    # random-init weights give an arbitrary split: shift the Stage-1 head bias so ~30 % are forwarded
    p = np.clip(_pipeline(128).run_waveform(rec, 48000).s1_probs.astype(np.float64), 1e-12, 1.0)
    bias = -float(np.quantile(np.log(p[:, 1]) - np.log(p[:, 0]), 0.7))
    pipe = _pipeline(128, head_bias1=bias)
    return rec, pipe, pipe.run_waveform(rec, 48000), bias


def test_cfg2_geometry_and_gate_bit_exact(cfg2):
    from oracle import glue

    rec, pipe, res, _ = cfg2
    assert rec.shape[-1] == 28_800_000 and res.num_windows == 1199  # 9.6 M samples -> 1199 windows
    assert res.s1_probs.shape == (1199, 2) and res.s1_probs.dtype == np.float32
    assert np.all(np.abs(res.s1_probs.sum(axis=1) - 1.0) <= 1e-6)
    preds, idx = glue.stage1_gate(res.s1_probs, np.float32(pipe.thr1))  # ref:312-320 on OUR probabilities
    assert np.array_equal(preds, res.s1_preds)
    assert np.array_equal(idx, res.swallow_indices) and idx.dtype == np.int64
    assert np.all(np.diff(res.swallow_indices) > 0)
    assert 200 < len(idx) < 600, "the calibrated head bias must forward roughly 30 % of the windows"
    cls = glue.stage2_classes(1199, res.stage2_results, np.float32(pipe.thr2), False)  # ref:332-340
    assert np.array_equal(cls, res.classes)
    doc = glue.summarize_stage_outputs(res.s1_probs, res.stage2_results, np.float32(pipe.thr2), False)  # ref:148-195
    for k, v in doc.items():
        if isinstance(v, float) and np.isnan(v):
            assert np.isnan(res.summary[k])
        else:
            assert res.summary[k] == v, k


def test_cfg2_independent_of_batch_composition(cfg2):
    """Per-window arithmetic must not depend on which windows share a launch: bit-identical scores."""
    rec, _, res, bias = cfg2
    other = _pipeline(37, head_bias1=bias).run_waveform(rec, 48000)
    assert np.array_equal(other.s1_probs, res.s1_probs)
    assert np.array_equal(other.swallow_indices, res.swallow_indices)
    assert np.array_equal(other.s2_probs, res.s2_probs)


def test_cfg2_prefix_locality(cfg2):
    """The first minute of the recording on its own reproduces the first windows of the full run bit-exactly (the
    resampler's 41-tap support only reaches the last 7 output samples of the prefix, i.e. its last window)."""
    rec, pipe, res, _ = cfg2
    part = pipe.run_waveform(rec[..., : 60 * 48000], 48000)
    assert part.num_windows == 119
    assert np.array_equal(part.s1_probs[:118], res.s1_probs[:118])
    assert np.abs(part.s1_probs[118] - res.s1_probs[118]).max() <= 1e-3


def test_cfg2_stage2_equals_direct_forward(cfg2):
    """Stage 2 on the compacted index list == the Stage-2 model applied to exactly those windows through the
    ASTFeatureExtractor / forward contract (gather, normalisation and compaction cannot have mixed windows up)."""
    from zenker_audio_detection_b200 import ops

    rec, pipe, res, _ = cfg2
    audio = ops.resample(torch.from_numpy(rec).cuda(), 48000, 16000)
    pick = res.swallow_indices[:: max(1, len(res.swallow_indices) // 24)][:24]
    wins = torch.stack([audio[int(k) * 8000: int(k) * 8000 + 16000] for k in pick])
    feats = pipe.fx2(wins, sampling_rate=16000, return_tensors="pt")["input_values"]
    probs = torch.softmax(pipe.m2(feats.to(pipe.device)).logits, dim=1).cpu().numpy()
    sel = np.searchsorted(res.swallow_indices, pick)
    assert np.abs(probs - res.s2_probs[sel]).max() <= 2e-3  # same kernels; the fused path normalises on the fly


@pytest.mark.parametrize("fixture,tag", [("cascade_cfg2.npz", "a"), ("cascade_cfg2.npz", "b"),
                                         ("cascade_300s_seed4242.npz", "a"), ("cascade_300s_seed4242.npz", "b")])
def test_cfg2_decisions_equal_the_reference_run(golden_dir, fixture, tag):
    """The headline configuration against the reference ITSELF: tests/golden/cascade_cfg2.npz holds what
    ref.window_audio / ref.forward_probs / the reference gate / ref.summarize_stage_outputs produced on the CPU for this
    very 600-s recording and these weights (scripts/make_golden.py::gold_cascade_cfg2, ~25 min of CPU), at thresholds
    0.5 / 0.5 (a) and 0.6 / 0.35 (b, the counting quirk of SURVEY.md 0.7).  Every integer of the result must be equal --
    the forwarded index list over all 1199 windows, the per-window classes, every count and ratio of the summary -- up
    to the one window of this recording whose reference margin (4.8e-6) is inside fp32 platform noise (see TAU below).
    cascade_300s_seed4242.npz is the same for an INDEPENDENT 300-s recording (another generator seed, 599 windows,
    thresholds 0.5 / 0.5 and 0.55 / 0.45; scripts/make_golden.py::gold_cascade_second)."""
    import json
    import os

    from oracle import glue
    from zenker_audio_detection_b200 import synth
    from zenker_audio_detection_b200.fx import ZenkerASTFeatureExtractor
    from zenker_audio_detection_b200.model import ZenkerASTForAudioClassification
    from zenker_audio_detection_b200.pipeline import TwoStagePipeline

    g = np.load(os.path.join(golden_dir, fixture))
    seconds = float(g["seconds"]) if "seconds" in g.files else 600.0
    seed = int(g["seed"]) if "seed" in g.files else 2002
    NW = int(g["n_windows"])
    thr1, thr2 = (float(v) for v in g[f"thresholds_{tag}"])
    fx1 = ZenkerASTFeatureExtractor(mean=synth.STAGE1_MEAN, std=synth.STAGE1_STD)
    fx2 = ZenkerASTFeatureExtractor(mean=synth.STAGE2_MEAN, std=synth.STAGE2_STD)
    m1 = ZenkerASTForAudioClassification({"max_length": 1024}, synth.random_state_dict(11, head_bias1=float(g["head_bias1_s1"])))
    m2 = ZenkerASTForAudioClassification({"max_length": 1024}, synth.random_state_dict(22, head_bias1=float(g["head_bias1_s2"])))
    pipe = TwoStagePipeline(m1, fx1, m2, fx2, batch_size=128, stage1_threshold=thr1, stage2_threshold=thr2)
    res = pipe.run_waveform(synth.recording(seconds, 48000, seed=seed), 48000)
    assert res.num_windows == NW == int((seconds - 1.0) / 0.5) + 1
    ref_idx, ref_s2 = g[f"swallow_indices_{tag}"], g[f"s2_probs_{tag}"]
    p = g["s1_probs"].astype(np.float64)
    margin = np.log(p[:, 1]) - np.log(p[:, 0])
    print(f"{fixture}[{tag}] thr {thr1}/{thr2}: re-checked {res.rechecked_s1} + {res.rechecked_s2} windows; forwarded ours/ref "
          f"{len(res.swallow_indices)}/{len(ref_idx)}; max |p1 - ref| {np.abs(res.s1_probs - g['s1_probs']).max():.3g}; "
          f"smallest reference |margin - decision point| "
          f"{min(np.abs(margin - m).min() for m in pipe.margins1):.3g}")
    # fp32 results move by a few 1e-6 in the logits between platforms (torch's own fp32 forward on this GPU is 3e-6 ..
    # 4e-6 from the CPU logits, tests/test_gpu_model.py), so a reference margin below TAU is not a decision any other
    # fp32 evaluation is bound to reproduce -- this recording has ONE such window (4.8e-6 from the argmax point; the next
    # is at 2.4e-5).  Every other window must decide exactly as the reference did.
    TAU = 1e-5
    ours_mask, ref_mask = np.zeros(NW, bool), np.zeros(NW, bool)
    ours_mask[res.swallow_indices] = True
    ref_mask[ref_idx] = True
    undecidable = np.zeros(NW, bool)
    for mg in pipe.margins1:
        undecidable |= np.abs(margin - mg) < TAU
    flips = ours_mask != ref_mask
    print(f"  windows within {TAU} of a Stage-1 decision point: {np.where(undecidable)[0].tolist()} "
          f"(margins {margin[undecidable].tolist()}); gate flips: {np.where(flips)[0].tolist()}")
    assert int(undecidable.sum()) <= 3  # cfg2 has one such window, the 300-s recording two (neither of which flips)
    assert not (flips & ~undecidable).any()
    assert np.abs(res.s1_probs - g["s1_probs"]).max() <= 2.5e-3
    both = ours_mask & ref_mask
    ours2 = res.s2_probs[np.searchsorted(res.swallow_indices, np.where(both)[0])]
    ref2 = ref_s2[np.searchsorted(ref_idx, np.where(both)[0])]
    assert np.abs(ours2 - ref2).max() <= 2.5e-3
    ref_classes = glue.stage2_classes(NW, [(int(i), q) for i, q in zip(ref_idx, ref_s2)], np.float32(thr2))
    assert np.array_equal(res.classes[~undecidable], ref_classes[~undecidable])
    ref_summary = json.loads(str(g[f"summary_{tag}"]))
    slack = int(undecidable.sum())  # each undecidable window may move one count by one
    for k, v in ref_summary.items():
        if isinstance(v, int):
            assert abs(res.summary[k] - v) <= slack, k
        elif v is None:
            assert res.summary[k] is None, k
        elif isinstance(v, float):
            assert abs(res.summary[k] - v) <= 2.5e-3 + slack / 17.0, k
        else:
            assert np.abs(np.asarray(res.summary[k]) - np.asarray(v)).max() <= 2.5e-3 + slack / 17.0, k


@pytest.fixture(scope="module")
def cfg3():
    from zenker_audio_detection_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(3003)
    wave = torch.randn(57_600_000, device="cuda", generator=g) * 0.05
    plan = ops.FbankPlan()
    return wave, plan, plan.fbank(wave)


def test_cfg3_shape_and_frame_independence(cfg3):
    wave, plan, fb = cfg3
    assert fb.shape == (359_998, 128)  # 1 + (57.6 M - 400) // 160
    assert bool(torch.isfinite(fb).all())
    shifted = plan.fbank(wave[160 * 1000:160 * 1000 + 160 * 5000 + 240])  # frames 1000 .. 5999 of the full signal
    assert shifted.shape[0] == 5000
    assert torch.equal(shifted, fb[1000:6000])


def test_cfg3_contract_rows_equal_strided_rows(cfg3):
    """SURVEY.md 0.9 at scale: fx(window k)[0:98] == (continuous fbank rows [50 k, 50 k + 98) - mean) / (2 std)."""
    wave, plan, fb = cfg3
    ks = torch.tensor([0, 1, 77, 3599, 7198], device="cuda")  # 7199 windows in one hour
    wins = torch.stack([wave[int(k) * 8000: int(k) * 8000 + 16000] for k in ks])
    feats = plan.fx_contract(wins, -1.15, 3.53, 1024, True)
    for i, k in enumerate(ks.tolist()):
        rows = (fb[50 * k: 50 * k + 98] - torch.tensor(-1.15, device="cuda")) / (torch.tensor(3.53, device="cuda") * 2)
        assert torch.equal(feats[i, :98], rows)
        assert torch.all(feats[i, 98:] == feats[i, 98, 0])  # pad rows hold the constant (0 - mean) / (2 std)


def test_cfg3_sampled_frames_vs_torchaudio(cfg3):
    from oracle import thirdparty

    wave, _, fb = cfg3
    for start in (0, 123_457, 359_000):
        piece = wave[start * 160: start * 160 + 160 * 199 + 400].cpu().numpy()
        ref = thirdparty.kaldi_fbank(piece)
        got = fb[start: start + 200].cpu().numpy()
        d = np.abs(got - ref)
        assert np.all(d <= 1e-4 * np.abs(ref) + 1e-3)
        assert np.mean(d <= 1e-4 * np.abs(ref)) >= 0.999


def test_cfg4_results_do_not_depend_on_order_or_rank():
    """Three recordings processed in two different orders / shard layouts give identical records after the gather
    bookkeeping (dist.pack_records / unpack_records); this is what makes the result independent of the world size."""
    from zenker_audio_detection_b200 import dist as zdist, synth

    pipe = _pipeline(64)
    recs = [synth.recording(30.0 + 7 * i, 48000, seed=4000 + i) for i in range(3)]
    shards = zdist.shard_recordings([r.shape[-1] for r in recs], 2)
    assert sorted(i for s in shards for i in s) == [0, 1, 2]

    def run(order):
        blocks = []
        for i in order:
            r = pipe.run_waveform(recs[i], 48000)
            blocks.append(zdist.pack_records(i, r.s1_probs, r.swallow_indices, r.s2_probs))
        return zdist.unpack_records(np.concatenate(blocks, axis=0))

    a, b = run([0, 1, 2]), run([i for s in reversed(shards) for i in s])
    assert a.keys() == b.keys()
    for rid in a:
        for x, y in zip(a[rid], b[rid]):
            assert np.array_equal(x, y)


def test_cfg4_a_recording_split_into_window_ranges_gives_the_same_records():
    """dist.shard_window_ranges cuts the pool into equal runs of windows, so a recording can be classified in two (or
    more) pieces on different ranks: `run_waveform(..., window_range=(w0, w1))` must give those windows BIT for bit what
    the unsplit run gives (same samples -> same frames -> same features; batch composition does not matter)."""
    from zenker_audio_detection_b200 import dist as zdist, synth
    from zenker_audio_detection_b200._lib import ZkError

    pipe = _pipeline(64)
    rec = synth.recording(47.0, 48000, seed=4321)
    whole = pipe.run_waveform(rec, 48000)
    n = whole.num_windows
    assert n == 93
    ref = zdist.pack_records(5, whole.s1_probs, whole.swallow_indices, whole.s2_probs)
    for cuts in ([0, 31, n], [0, 1, 64, 65, n], [0, n]):
        blocks = []
        for w0, w1 in zip(cuts[:-1], cuts[1:]):
            r = pipe.run_waveform(rec, 48000, window_range=(w0, w1))
            assert r.num_windows == w1 - w0
            blocks.append(zdist.pack_records(5, r.s1_probs, r.swallow_indices, r.s2_probs, window_base=w0))
        assert np.array_equal(np.concatenate(blocks), ref), cuts
    with pytest.raises(ZkError):
        pipe.run_waveform(rec, 48000, window_range=(10, n + 1))


def test_cfg5_batch32_forward_is_batch_invariant_and_matches_the_fp32_oracle():
    """cfg5: (32, 1024, 128) features ~ N(0, 0.5), seed 5005, 1214 tokens, 16-bit weights.  The
    32-window forward equals four 8-window forwards and thirty-two single-window forwards bit for bit (tile shapes do
    not depend on the batch), and agrees with the fp32 oracle (HF's arithmetic restated, oracle/numerics.py) within the
    bf16 tolerance of north_star on a plain random init."""
    from oracle import numerics
    from zenker_audio_detection_b200 import ops, synth

    g = torch.Generator(device="cuda").manual_seed(5005)
    feats = torch.randn(32, 1024, 128, device="cuda", generator=g) * 0.5
    sd = synth.random_state_dict(31, qk_gain=1.0)
    m = ops.AstModel(sd)
    full = m.forward_features(feats)
    assert full.shape == (32, 2) and full.dtype == torch.float32 and bool(torch.isfinite(full).all())
    by8 = torch.cat([m.forward_features(feats[i:i + 8]) for i in range(0, 32, 8)])
    by1 = torch.cat([m.forward_features(feats[i:i + 1]) for i in range(32)])
    assert torch.equal(full, by8) and torch.equal(full, by1)
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.inference_mode():
        ref = numerics.ast_forward({k: v.cuda() for k, v in sd.items()}, feats[:4])
    assert (full[:4] - ref).abs().max().item() <= 2e-3  # fp16 operands (1e-2 is north_star's bf16 bound)
