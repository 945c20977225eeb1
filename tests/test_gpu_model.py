"""Full AST forward through the C ABI vs the fp32 oracle (numerics.ast_forward) and the HF golden logits."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _oracle_logits(sd, feats, **kw):
    from oracle import numerics

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sdc = {k: v.cuda() for k, v in sd.items()}
    with torch.inference_mode():
        return numerics.ast_forward(sdc, feats.cuda(), **kw)


def test_one_layer_hidden_state():
    """1-layer model: the residual stream must match fp32 closely (pins token order, patch gather, pos-emb)."""
    from zenker_audio_detection_b200 import ops, synth

    sd = synth.random_state_dict(5)
    plan = ops.FbankPlan()
    w = torch.from_numpy(synth.cfg1_windows(4)).cuda()
    feats = plan.fx_contract(w, synth.STAGE1_MEAN, synth.STAGE1_STD, 1024)
    m = ops.AstModel(sd, num_layers=1)
    logits, hidden = m.forward_features(feats, return_hidden=True)
    ref_logits, ref_hidden = _oracle_logits(sd, feats, num_layers=1, return_hidden=True)
    diff = (hidden - ref_hidden)
    rel = (diff.norm() / ref_hidden.norm()).item()
    assert rel <= 1.2e-2, rel                    # bf16 operands (2^-9 each), fp32 accumulation / residual stream
    assert diff.abs().max().item() <= 8e-2       # max over 3.7 M elements
    assert (hidden[:, :2] - ref_hidden[:, :2]).abs().max().item() <= 3e-2  # cls / dist tokens
    assert (logits - ref_logits).abs().max().item() <= 2.5e-2  # sensitised init; the 1e-2 gate is the plain-init test below


def test_full_forward_plain_init_within_1e_2(golden_dir):
    """north_star: logits within 1e-2 absolute in bf16 -- on a random init at HF's own scale (query/key gain 1)."""
    from zenker_audio_detection_b200 import ops, synth

    gold = np.load(os.path.join(golden_dir, "ast_plain.npz"))
    plan = ops.FbankPlan()
    w = torch.from_numpy(synth.cfg1_windows(64)[:8]).cuda()
    sd = synth.random_state_dict(int(gold["seed"]), qk_gain=float(gold["qk_gain"]))
    feats = plan.fx_contract(w, synth.STAGE1_MEAN, synth.STAGE1_STD, 1024)
    logits = ops.AstModel(sd).forward_features(feats)
    ref = _oracle_logits(sd, feats)
    err = (logits - ref).abs().max().item()
    gerr = np.abs(logits.cpu().numpy() - gold["logits"]).max()
    print(f"plain init: max |logit - fp32 oracle| = {err:.4g}; vs HF golden = {gerr:.4g}")
    assert err <= 1e-2 and gerr <= 1e-2, (err, gerr)


def test_full_forward_sensitised_init_decisions(golden_dir):
    """The conditioned init (query/key x4) amplifies bf16 noise (CPU bf16 autocast: 1.2e-2 ..
    1.9e-2); logits must stay within 2.5e-2 and every thresholded decision whose fp32 margin exceeds the measured
    error must be identical."""
    from zenker_audio_detection_b200 import ops, synth

    gold = np.load(os.path.join(golden_dir, "ast_cfg1.npz"))
    plan = ops.FbankPlan()
    w = torch.from_numpy(synth.cfg1_windows(64)[:16]).cuda()
    for seed, (mean, std), key, bkey in ((11, (synth.STAGE1_MEAN, synth.STAGE1_STD), "logits1", "head_bias1_s1"),
                                         (22, (synth.STAGE2_MEAN, synth.STAGE2_STD), "logits2", "head_bias1_s2")):
        sd = synth.random_state_dict(seed, head_bias1=float(gold[bkey]))
        feats = plan.fx_contract(w, mean, std, 1024)
        m = ops.AstModel(sd)
        logits = m.forward_features(feats)
        ref = _oracle_logits(sd, feats).cpu().numpy()
        got = logits.cpu().numpy()
        err = np.abs(got - ref).max()
        gerr = np.abs(got - gold[key]).max()
        assert np.abs(ref - gold[key]).max() <= 1e-4  # fp32 oracle restatement == HF on the CPU (golden)
        margin = gold[key][:, 1] - gold[key][:, 0]
        band = np.abs(margin) <= 2 * gerr
        flips = ((got[:, 1] > got[:, 0]) != (margin > 0)) & ~band
        print(f"seed {seed}: max |logit - fp32| = {err:.4g} (golden {gerr:.4g}); min |margin| = {np.abs(margin).min():.4g}; "
              f"in band {int(band.sum())}; flips outside band {int(flips.sum())}")
        assert gerr <= 2.5e-2, gerr
        assert not flips.any()
        del m


def test_fused_fbank_path_matches_contract_path():
    from zenker_audio_detection_b200 import ops, synth

    sd = synth.random_state_dict(11)
    plan = ops.FbankPlan()
    wave = torch.from_numpy(synth.recording(8.0, 16000, seed=3)).cuda()
    fb = plan.fbank(wave)
    nwin = (wave.numel() - 16000) // 8000 + 1
    wins = torch.stack([wave[8000 * k: 8000 * k + 16000] for k in range(nwin)])
    feats = plan.fx_contract(wins, synth.STAGE1_MEAN, synth.STAGE1_STD, 1024)
    m = ops.AstModel(sd, num_layers=2)
    a = m.forward_features(feats)
    b = m.forward_fbank(fb, nwin, synth.STAGE1_MEAN, synth.STAGE1_STD)
    assert torch.equal(a, b)
    idx = torch.tensor([5, 2, 9], dtype=torch.int32, device="cuda")
    c = m.forward_fbank(fb, 3, synth.STAGE1_MEAN, synth.STAGE1_STD, window_index=idx)
    assert torch.equal(c, a[idx.long()])


@pytest.mark.parametrize("batch", [1, 5])
def test_last_layer_tail_matches_full_layer(batch):
    """The pruned last layer (K/V for every token, everything else for tokens 0/1 only) must give the logits of the
    full layer: same inputs, same bf16 operands; only the 2-query attention runs in fp32 instead of bf16 P."""
    from zenker_audio_detection_b200 import ops, synth

    sd = synth.random_state_dict(9)
    plan = ops.FbankPlan()
    w = torch.from_numpy(synth.cfg1_windows(8)[:batch]).cuda()
    feats = plan.fx_contract(w, synth.STAGE1_MEAN, synth.STAGE1_STD, 1024)
    for layers in (1, 12):
        m = ops.AstModel(sd, num_layers=layers)
        pruned = m.forward_features(feats)
        full, _ = m.forward_features(feats, return_hidden=True)  # asking for the hidden state forces the full layer
        err = (pruned - full).abs().max().item()
        assert err <= 4e-3, (layers, err)
