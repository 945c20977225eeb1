"""LayerNorm, gate/compaction, softmax through the C ABI vs torch / the oracle glue."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_layernorm():
    from zenker_audio_detection_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(1214 * 2 + 3, 768, device="cuda", generator=g) * 3 + 0.5
    w = torch.randn(768, device="cuda", generator=g)
    b = torch.randn(768, device="cuda", generator=g)
    out = ops.layernorm(x, w, b, 1e-12)
    ref = torch.nn.functional.layer_norm(x, (768,), w, b, 1e-12)
    err = (out.float() - ref).abs().max().item()
    assert err <= 8e-3 * ref.abs().max().item(), err
    assert torch.equal(out, ref.to(torch.bfloat16)) or (out.float() - ref.to(torch.bfloat16).float()).abs().max() <= 0.04


def test_layernorm_fp16_and_planes():
    from zenker_audio_detection_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(1214 + 3, 768, device="cuda", generator=g) * 3 + 0.5
    w = torch.randn(768, device="cuda", generator=g)
    b = torch.randn(768, device="cuda", generator=g)
    ref = torch.nn.functional.layer_norm(x.double(), (768,), w.double(), b.double(), 1e-12)
    one = ops.layernorm(x, w, b, 1e-12, dtype=torch.float16)
    assert one.dtype == torch.float16 and (one.double() - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()
    two = ops.layernorm(x, w, b, 1e-12, dtype=torch.float16, planes=2)
    assert two.shape == (x.shape[0], 1536)
    assert torch.equal(two[:, :768], one)  # the hi plane IS the single-plane result
    val = two[:, :768].double() + two[:, 768:].double()
    assert (val - ref).abs().max().item() <= 4e-6 * ref.abs().max().item()  # fp32 LayerNorm arithmetic


@pytest.mark.parametrize("n", [0, 1, 33, 1199, 5000])
def test_band_select_and_scatter(n):
    from zenker_audio_detection_b200 import ops

    g = torch.Generator().manual_seed(n + 7)
    logits = (torch.randn(n, 2, generator=g) * 0.2)
    margins, eps = [0.0, 0.405], 0.03
    d = (logits[:, 1] - logits[:, 0]).numpy()
    if n > 3:
        logits[2, 1] = float("nan")  # a NaN margin must be selected (never silently trusted)
        d = (logits[:, 1] - logits[:, 0]).numpy()
    want = np.where(~(np.minimum(np.abs(d - np.float32(0.0)), np.abs(d - np.float32(0.405))) > np.float32(eps)))[0]
    src = torch.arange(n, dtype=torch.int32) * 3 + 1
    dl = logits.cuda()
    pos, window, count = ops.band_select(dl, margins, eps, src.cuda())
    r = int(count.item())
    assert r == len(want)
    assert np.array_equal(pos[:r].cpu().numpy(), want.astype(np.int32))
    assert np.array_equal(window[:r].cpu().numpy(), (want * 3 + 1).astype(np.int32))
    pos2, window2, count2 = ops.band_select(dl, margins, eps)
    assert int(count2.item()) == r and torch.equal(pos2[:r], window2[:r])
    if r:
        new = torch.arange(2 * r, dtype=torch.float32).view(r, 2).cuda() + 100
        before = dl.clone()
        ops.scatter_rows2(new, pos, r, dl)
        keep = np.ones(n, bool)
        keep[want] = False
        assert torch.equal(dl[torch.from_numpy(want).cuda()], new)
        assert torch.equal(dl[torch.from_numpy(np.where(keep)[0]).cuda()], before[torch.from_numpy(np.where(keep)[0]).cuda()])


@pytest.mark.parametrize("n", [0, 1, 31, 1024, 1199, 7199, 50000])
@pytest.mark.parametrize("thr,minp", [(0.5, None), (0.8, None), (0.5, 0.7), (0.3, None)])
def test_gate_compact_bit_exact(n, thr, minp):
    from oracle import glue
    from zenker_audio_detection_b200 import ops

    g = torch.Generator().manual_seed(n + 1)
    logits = (torch.randn(n, 2, generator=g) * 1.5).cuda()
    if n > 10:
        logits[3] = logits[3, 0]  # exact tie -> argmax picks class 0
        logits[5, 1] = logits[5, 0] + 1e-7
    probs, pred, index, count = ops.gate_compact(logits, thr, minp)
    k = int(count.item())
    p = probs.cpu().numpy()
    # integer outputs must be bit-exact functions of the probabilities (oracle: ref:312-320 / refc:471-478)
    opred, oidx = glue.stage1_gate(p, np.float32(thr), None if minp is None else np.float32(minp))
    assert np.array_equal(pred.cpu().numpy(), opred.astype(np.int32))
    assert k == len(oidx)
    assert np.array_equal(index[:k].cpu().numpy(), oidx.astype(np.int32))
    # and the probabilities are torch.softmax's up to rounding
    if n:
        ref = torch.softmax(logits, dim=1).cpu().numpy()
        assert np.abs(p - ref).max() <= 2e-7


def test_softmax2():
    from zenker_audio_detection_b200 import ops

    logits = torch.randn(777, 2, device="cuda") * 4
    p = ops.softmax2(logits)
    assert (p - torch.softmax(logits, dim=1)).abs().max().item() <= 2e-7


@pytest.mark.parametrize("n", [1, 3, 1027, 98 * 128 * 64])
def test_feature_stats_match_the_float64_reference(n):
    """utils/compute_ast_normalization_stats.py:77-95: float64 sums over the zero-padded features, unbiased std."""
    from zenker_audio_detection_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(n)
    x = torch.randn(n, device="cuda", generator=g) * 3.5 - 4.2
    st = ops.FeatureStats()
    half = n // 2
    padded = 2 * n + 5  # the values stand for a zero-padded tensor of this many elements
    st.update(x[:half].contiguous() if half % 4 == 0 else x[:half].clone(), padded_elements=padded - (n - half))
    st.update(x[half:].clone(), padded_elements=n - half)
    flat = torch.cat([x.double().cpu(), torch.zeros(padded - n, dtype=torch.float64)])
    mean = flat.sum().item() / padded
    var = max((flat ** 2).sum().item() / padded - mean * mean, 0.0) * (padded / (padded - 1))
    got = st.result()
    assert got["count"] == padded
    assert abs(got["mean"] - mean) <= 1e-12 * max(1.0, abs(mean)) + 1e-13
    assert abs(got["std"] - var ** 0.5) <= 1e-10 * max(1.0, var ** 0.5)
