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
