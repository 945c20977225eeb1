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


@pytest.mark.parametrize("fmt", ["fp16", "bf16"])
def test_one_layer_hidden_state(fmt):
    """1-layer model: the residual stream must match fp32 closely (pins token order, patch gather, pos-emb)."""
    from zenker_audio_detection_b200 import ops, synth

    sd = synth.random_state_dict(5)
    plan = ops.FbankPlan()
    w = torch.from_numpy(synth.cfg1_windows(4)).cuda()
    feats = plan.fx_contract(w, synth.STAGE1_MEAN, synth.STAGE1_STD, 1024)
    m = ops.AstModel(sd, num_layers=1, operand_format=fmt)
    logits, hidden = m.forward_features(feats, return_hidden=True)
    ref_logits, ref_hidden = _oracle_logits(sd, feats, num_layers=1, return_hidden=True)
    diff = (hidden - ref_hidden)
    rel = (diff.norm() / ref_hidden.norm()).item()
    k = 1.0 if fmt == "bf16" else 0.2            # fp16 operands carry 3 more bits (2^-12 against 2^-9)
    assert rel <= 1.2e-2 * k, rel                # 16-bit operands, fp32 accumulation / residual stream
    assert diff.abs().max().item() <= 8e-2 * k   # max over 3.7 M elements
    assert (hidden[:, :2] - ref_hidden[:, :2]).abs().max().item() <= 3e-2 * k  # cls / dist tokens
    assert (logits - ref_logits).abs().max().item() <= 2.5e-2 * k


def test_one_layer_hidden_state_recheck_precision():
    """Split-operand forward, one layer: the whole residual stream is fp32-class (every kernel of the re-check path:
    plane gather, split patch GEMM, LayerNorm planes, split QKV / attention / out-proj / MLP with erf GELU)."""
    from zenker_audio_detection_b200 import _lib, ops, synth

    sd = synth.random_state_dict(5)
    plan = ops.FbankPlan()
    w = torch.from_numpy(synth.cfg1_windows(3)).cuda()
    feats = plan.fx_contract(w, synth.STAGE1_MEAN, synth.STAGE1_STD, 1024)
    m = ops.AstModel(sd, num_layers=1)
    logits, hidden = m.forward_features(feats, return_hidden=True, precision=_lib.PRECISION_RECHECK)
    ref_logits, ref_hidden = _oracle_logits(sd, feats, num_layers=1, return_hidden=True)
    diff = (hidden - ref_hidden)
    rel = (diff.norm() / ref_hidden.norm()).item()
    print(f"recheck precision, 1 layer: rel {rel:.3e}, max {diff.abs().max().item():.3e}, "
          f"logits {(logits - ref_logits).abs().max().item():.3e}")
    # two fp32-class evaluations of K = 768 .. 3072 dot products with cancellation differ by ~1e-5 relative (measured
    # 7.7e-6; torch's own fp32 GEMM is that far from float64, tests/test_gpu_gemm.py::test_split_gemm_matches_float64)
    assert rel <= 2e-5, rel
    assert diff.abs().max().item() <= 6e-5
    assert (logits - ref_logits).abs().max().item() <= 2.5e-5


@pytest.mark.parametrize("fmt", ["fp16", "bf16"])
def test_full_forward_plain_init_within_1e_2(golden_dir, fmt):
    """north_star: logits within 1e-2 absolute in bf16 -- on a random init at HF's own scale (query/key gain 1)."""
    from zenker_audio_detection_b200 import ops, synth

    gold = np.load(os.path.join(golden_dir, "ast_plain.npz"))
    plan = ops.FbankPlan()
    w = torch.from_numpy(synth.cfg1_windows(64)[:8]).cuda()
    sd = synth.random_state_dict(int(gold["seed"]), qk_gain=float(gold["qk_gain"]))
    feats = plan.fx_contract(w, synth.STAGE1_MEAN, synth.STAGE1_STD, 1024)
    logits = ops.AstModel(sd, operand_format=fmt).forward_features(feats)
    ref = _oracle_logits(sd, feats)
    err = (logits - ref).abs().max().item()
    gerr = np.abs(logits.cpu().numpy() - gold["logits"]).max()
    print(f"plain init {fmt}: max |logit - fp32 oracle| = {err:.4g}; vs HF golden = {gerr:.4g}")
    tol = 1e-2 if fmt == "bf16" else 2e-3
    assert err <= tol and gerr <= tol, (err, gerr)


def _conditioned_cases(gold):
    from zenker_audio_detection_b200 import synth

    return ((11, (synth.STAGE1_MEAN, synth.STAGE1_STD), "logits1", "head_bias1_s1"),
            (22, (synth.STAGE2_MEAN, synth.STAGE2_STD), "logits2", "head_bias1_s2"))


def test_full_forward_sensitised_init_within_1e_2(golden_dir):
    """The conditioned init (query/key x4) is what cfg1 / cfg2 / the bench run on.  The FAST
    path (fp16 operands, the default) must hold north_star's 1e-2 logit bound on it against the HF CPU logits; bf16
    operands, kept selectable, are 5-8x noisier (measured 1.6e-2 / 1.85e-2) and are only held to 2.5e-2."""
    from zenker_audio_detection_b200 import ops, synth

    gold = np.load(os.path.join(golden_dir, "ast_cfg1.npz"))
    plan = ops.FbankPlan()
    w = torch.from_numpy(synth.cfg1_windows(64)[:16]).cuda()
    for seed, (mean, std), key, bkey in _conditioned_cases(gold):
        sd = synth.random_state_dict(seed, head_bias1=float(gold[bkey]))
        feats = plan.fx_contract(w, mean, std, 1024)
        ref = _oracle_logits(sd, feats).cpu().numpy()
        assert np.abs(ref - gold[key]).max() <= 1e-4  # fp32 oracle restatement == HF on the CPU (golden)
        for fmt, tol in (("fp16", 1e-2), ("bf16", 2.5e-2)):
            m = ops.AstModel(sd, operand_format=fmt)
            got = m.forward_features(feats).cpu().numpy()
            gerr = np.abs(got - gold[key]).max()
            print(f"seed {seed} {fmt}: max |logit - HF golden| = {gerr:.4g}")
            assert gerr <= tol, (fmt, gerr)
            del m


def test_recheck_precision_is_fp32_class_and_decisions_identical(golden_dir):
    """ZK_PRECISION_RECHECK on the conditioned weights: logits within 4e-5 of the HF CPU fp32 logits (measured 1.4e-5 /
    1.9e-5; torch's fp32 forward on the same GPU is itself printed for comparison -- fp32 results move by ~1e-5 between
    platforms), and after the in-place re-check of the borderline windows EVERY thresholded decision (argmax, and
    p1 >= 0.6 as a second decision point) equals the reference's -- no exemption band."""
    from zenker_audio_detection_b200 import _lib, ops, synth
    from zenker_audio_detection_b200.model import decision_margins

    gold = np.load(os.path.join(golden_dir, "ast_cfg1.npz"))
    plan = ops.FbankPlan()
    w = torch.from_numpy(synth.cfg1_windows(64)[:16]).cuda()
    for seed, (mean, std), key, bkey in _conditioned_cases(gold):
        sd = synth.random_state_dict(seed, head_bias1=float(gold[bkey]))
        feats = plan.fx_contract(w, mean, std, 1024)
        m = ops.AstModel(sd)
        hi = m.forward_features(feats, precision=_lib.PRECISION_RECHECK).cpu().numpy()
        err = np.abs(hi - gold[key]).max()
        gpu32 = np.abs(_oracle_logits(sd, feats).cpu().numpy() - gold[key]).max()
        print(f"seed {seed}: recheck precision max |logit - HF CPU golden| = {err:.3e}; torch fp32 on this GPU vs the same "
              f"golden = {gpu32:.3e}")
        assert err <= 4e-5, err
        # the product flow: fast logits, then re-check with a band wide enough to catch several of the 16 windows
        ref_margin = gold[key][:, 1] - gold[key][:, 0]
        margins = decision_margins([0.5, 0.6])
        fast = m.forward_features(feats)
        fast_err = np.abs(fast.cpu().numpy() - gold[key]).max()
        n = m.recheck_features(feats, fast, margins, eps=0.05)
        got = fast.cpu().numpy()
        d = got[:, 1] - got[:, 0]
        in_band = np.zeros(16, bool)
        for mg in margins:
            in_band |= np.abs(ref_margin - mg) < 0.05 - fast_err
        assert n >= int(in_band.sum())
        assert np.abs(got[in_band] - gold[key][in_band]).max() <= 4e-5 if in_band.any() else True
        for mg in margins:
            assert np.array_equal(d > mg, ref_margin > mg), (seed, mg)
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
    full layer: same inputs, same 16-bit operands; only the 2-query attention runs in fp32 instead of 16-bit P."""
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


@pytest.mark.parametrize("max_length", [256, 112])
def test_other_max_length_geometries(max_length):
    """ASTConfig.max_length is a checkpoint property (HF:configuration...: max_length, ref:91-93 passes the config
    through): a model fine-tuned on max_length 256 has 2 + 12 * 25 = 302 tokens, one on 112 has 122 -- fewer than one
    query tile, with ragged key blocks and row tiles in every kernel.  Full 3-layer forward at both precisions against
    the fp32 oracle, and through the fused fbank gather (frames beyond the window are the pad constant)."""
    from zenker_audio_detection_b200 import _lib, ops, synth

    tokens = 2 + 12 * ((max_length - 16) // 10 + 1)
    sd = synth.random_state_dict(31)
    g = torch.Generator().manual_seed(max_length)
    sd["audio_spectrogram_transformer.embeddings.position_embeddings"] = torch.randn(1, tokens, 768, generator=g) * 0.02
    plan = ops.FbankPlan()
    w = torch.from_numpy(synth.cfg1_windows(9)).cuda()
    feats = plan.fx_contract(w, synth.STAGE1_MEAN, synth.STAGE1_STD, max_length)
    assert feats.shape == (9, max_length, 128)
    m = ops.AstModel(sd, max_length=max_length, num_layers=3)
    assert m.tokens == tokens
    ref, ref_hidden = _oracle_logits(sd, feats, num_layers=3, return_hidden=True)
    fast, hidden = m.forward_features(feats, return_hidden=True)
    pruned = m.forward_features(feats)
    hi = m.forward_features(feats, precision=_lib.PRECISION_RECHECK)
    e_fast, e_pruned, e_hi = ((t - ref).abs().max().item() for t in (fast, pruned, hi))
    rel = ((hidden - ref_hidden).norm() / ref_hidden.norm()).item()
    print(f"max_length {max_length} ({tokens} tokens): fast {e_fast:.3g}, pruned {e_pruned:.3g}, re-check {e_hi:.3g}, hidden rel {rel:.3g}")
    assert e_fast <= 5e-3 and e_pruned <= 5e-3 and rel <= 3e-3, (e_fast, e_pruned, rel)
    assert e_hi <= 5e-5, e_hi


def test_layernorm_tail_of_the_residual_gemms_is_bit_identical():
    """ZK_LN_FUSE (north_star (4): LayerNorm fused into the GEMM that produces its input): bit 0 lets fc2 write the next
    layer's layernorm_before output, bit 1 lets the out-projection write layernorm_after, both from extra warps of the
    GEMM kernel while the rows are still in L2.  Same row arithmetic as the separate kernel, so logits and the whole
    residual stream must be BIT-identical in every mode -- with the last layer pruned (logits) and unpruned (hidden).
    80 windows = 380 row tiles, i.e. ~15 tiles per CTA pair, so the hand-over between CTAs is exercised across many
    tiles.  The switch is read once per process, hence the subprocesses."""
    import hashlib
    import subprocess
    import sys

    code = (
        "import torch, sys, hashlib; sys.path.insert(0, %r)\n"
        "from zenker_audio_detection_b200 import ops, synth\n"
        "plan = ops.FbankPlan()\n"
        "w = torch.from_numpy(synth.cfg1_windows(80)).cuda()\n"
        "feats = plan.fx_contract(w, synth.STAGE1_MEAN, synth.STAGE1_STD, 1024)\n"
        "m = ops.AstModel(synth.random_state_dict(5), num_layers=3)\n"
        "l1 = m.forward_features(feats)\n"
        "l2, hid = m.forward_features(feats, return_hidden=True)\n"
        "torch.cuda.synchronize()\n"
        "print('SUM', hashlib.sha256(l1.cpu().numpy().tobytes() + l2.cpu().numpy().tobytes() + hid.cpu().numpy().tobytes()).hexdigest(),"
        " float(l1.abs().sum()))\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sums = {}
    for mode in ("0", "1", "2", "3"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, ZK_LN_FUSE=mode), capture_output=True, text=True)
        assert r.returncode == 0, (mode, r.stdout[-400:], r.stderr[-1500:])
        sums[mode] = [ln for ln in r.stdout.splitlines() if ln.startswith("SUM")][0]
    print(sums)
    assert len(set(sums.values())) == 1, sums
