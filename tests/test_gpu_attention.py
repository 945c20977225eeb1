"""Fused tcgen05 attention vs torch fp32 softmax(QK^T/8)V on the same bf16 q,k,v (through the C ABI)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(qkv, B, T):
    q, k, v = (qkv[:, i * 768:(i + 1) * 768].float().view(B, T, 12, 64).transpose(1, 2) for i in range(3))
    s = (q @ k.transpose(2, 3)) * 0.125
    o = torch.softmax(s, dim=-1) @ v
    return o.transpose(1, 2).reshape(B * T, 768)


@pytest.mark.parametrize("B,T,scale", [(2, 1214, 1.0), (1, 128, 1.0), (1, 129, 1.0), (3, 300, 3.0), (1, 62, 1.0),
                                       (1, 1214, 6.0), (2, 192, 1.0), (1, 193, 2.0),  # last key block of exactly 64 / 65 keys (half-width path)
                                       (16, 1214, 2.0), (40, 200, 4.0)])  # > 148 work items: persistent CTAs walk several
def test_attention(B, T, scale):
    from zenker_audio_detection_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(B * 1000 + T)
    qkv = torch.randn(B * T, 2304, device="cuda", generator=g)
    qkv[:, :1536] *= scale  # larger logits -> peaky softmax
    qkv = qkv.to(torch.bfloat16)
    out = ops.attention(qkv, B, T)
    torch.cuda.synchronize()
    ref = _ref(qkv, B, T)
    err = (out.float() - ref).abs().max().item()
    # bf16 output: half an ulp at |o| in [4, 8) is 1.56e-2 on its own, plus the bf16 rounding of P (2^-9 relative)
    # against an fp32 row sum: 2^-9 * |v|max ~ 8e-3 for a one-hot row (scale 6 makes most rows one-hot)
    assert err <= 3.5e-2, err  # (one bf16 ulp at |o| in [4, 8) is 3.1e-2)
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    assert rel <= 1e-2, rel


@pytest.mark.parametrize("seed", [3, 19])
def test_attention_does_not_depend_on_the_neighbouring_window(seed):
    """The rows past the last token of a window (1214 = 9 * 128 + 62) belong to the NEXT window of the batch: they are
    masked as keys and never stored as queries, and they must not influence the window's own rows in any other way
    (seed 19 used to: an out-of-window row sharing a warp with real ones triggered the warp-wide max update)."""
    from zenker_audio_detection_b200 import ops

    T = 1214
    g = torch.Generator(device="cuda").manual_seed(seed)
    qkv = torch.randn(2 * T, 2304, device="cuda", generator=g)
    qkv[:, :1536] *= 2.5
    qkv = qkv.to(torch.bfloat16)
    both = ops.attention(qkv, 2, T)
    alone = ops.attention(qkv[:T].contiguous(), 1, T)
    assert torch.equal(both[:T], alone)


@pytest.mark.parametrize("B,T,scale", [(2, 1214, 1.0), (1, 129, 1.0), (3, 300, 3.0), (1, 1214, 6.0), (1, 193, 2.0)])
def test_attention_fp16(B, T, scale):
    from zenker_audio_detection_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(B * 1000 + T)
    qkv = torch.randn(B * T, 2304, device="cuda", generator=g)
    qkv[:, :1536] *= scale
    qkv = qkv.to(torch.float16)
    out = ops.attention(qkv, B, T)
    assert out.dtype == torch.float16
    ref = _ref(qkv, B, T)
    err = (out.float() - ref).abs().max().item()
    assert err <= 5e-3, err  # fp16 P (2^-12 relative) + fp16 output rounding (2^-12 of |o| <= 8) + the poly / MUFU exp2
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    assert rel <= 1.5e-3, rel


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("key", [256, 257, 258, 300, 391, 1213])
def test_attention_one_key_far_above_the_running_maximum(key, dtype):
    """A key in a LATER block whose score exceeds everything before it by hundreds of units: exp2 against the stale
    maximum overflows (MUFU.EX2 saturates to inf, the FMA-pipe polynomial wraps around), so the rescale trigger -- the
    block's row sum plus the maximum over the polynomial's own arguments -- must fire whichever slot the key falls
    into (256 / 257: polynomial pair, 258 / 300: MUFU pair, 391: odd position, 1213: the ragged half block)."""
    from zenker_audio_detection_b200 import ops

    T = 1214
    g = torch.Generator(device="cuda").manual_seed(key)
    qkv = torch.randn(T, 2304, device="cuda", generator=g) * 0.5
    qkv[:, :768] = qkv[:, :768].abs()                  # q >= 0 ...
    qkv[key, 768:1536] = 24.0                          # ... so this key scores ~ 24 * 64 * E|q| / 8 = 77 x the others
    qkv = qkv.to(dtype)
    out = ops.attention(qkv, 1, T)
    ref = _ref(qkv, 1, T)
    assert torch.isfinite(out.float()).all()
    err = (out.float() - ref).abs().max().item()
    assert err <= (5e-3 if dtype == torch.float16 else 3.5e-2), err
    # every row is (all but) one-hot on `key`
    v = qkv[key, 1536:].float()
    assert (out.float() - v).abs().max().item() <= 0.1


def _ref64(qkv32, B, T):
    q, k, v = (qkv32[:, i * 768:(i + 1) * 768].double().view(B, T, 12, 64).transpose(1, 2) for i in range(3))
    s = (q @ k.transpose(2, 3)) * 0.125
    o = torch.softmax(s, dim=-1) @ v
    return o.transpose(1, 2).reshape(B * T, 768)


@pytest.mark.parametrize("B,T,scale", [(2, 1214, 1.0), (1, 1214, 4.0), (1, 64, 1.0), (1, 65, 2.0), (3, 300, 3.0), (1, 7, 1.0),
                                       (5, 130, 6.0)])
def test_attention_split_matches_float64(B, T, scale):
    """Re-check precision: q, k, v as fp16 hi | lo planes, three-product contractions, exp2f softmax; against float64 on
    the fp32 inputs the result must be fp32-class (and the hi + lo output planes carry it)."""
    from zenker_audio_detection_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(B * 77 + T)
    qkv = torch.randn(B * T, 2304, device="cuda", generator=g)
    qkv[:, :1536] *= scale
    out = ops.attention_split(ops.split_f16(qkv), B, T)
    assert out.shape == (B * T, 1536)
    val = out[:, :768].double() + out[:, 768:].double()
    ref = _ref64(qkv, B, T)
    err = (val - ref).abs().max().item()
    # scores grow with scale^2 and their fp32-level error e^(delta s) - 1 with them, for ANY fp32 evaluation: the bar is
    # the error of torch's own fp32 attention on the same inputs (measured: 0.7-4x of it)
    err32 = (_ref(qkv, B, T).double() - ref).abs().max().item()
    print(f"split attention B={B} T={T} scale={scale}: err {err:.3e} (torch fp32: {err32:.3e})")
    assert err <= 5 * err32 + 2e-6, (err, err32)


def test_attention_split_window_independence():
    from zenker_audio_detection_b200 import ops

    T = 1214
    qkv = ops.split_f16(torch.randn(2 * T, 2304, device="cuda") * 2.0)
    both = ops.attention_split(qkv, 2, T)
    alone = ops.attention_split(qkv[:T].contiguous(), 1, T)
    assert torch.equal(both[:T], alone)


def test_attention_split_mma_variant_still_agrees():
    """ZK_SPLIT_ATTN=mma selects the warp-level mma.sync kernel instead of the tcgen05 one (the switch is read once per
    process, hence the subprocess): both must give float64-class results on the same inputs."""
    import os
    import subprocess
    import sys

    code = (
        "import torch, sys; sys.path.insert(0, %r)\n"
        "from zenker_audio_detection_b200 import ops\n"
        "g = torch.Generator(device='cuda').manual_seed(5)\n"
        "B, T = 2, 333\n"
        "qkv = torch.randn(B * T, 2304, device='cuda', generator=g); qkv[:, :1536] *= 2.0\n"
        "out = ops.attention_split(ops.split_f16(qkv), B, T)\n"
        "val = out[:, :768].double() + out[:, 768:].double()\n"
        "q, k, v = (qkv[:, i * 768:(i + 1) * 768].double().view(B, T, 12, 64).transpose(1, 2) for i in range(3))\n"
        "ref = (torch.softmax((q @ k.transpose(2, 3)) * 0.125, -1) @ v).transpose(1, 2).reshape(B * T, 768)\n"
        "err = (val - ref).abs().max().item(); print(err); assert err < 5e-6, err\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for variant in ("mma", "tc"):
        env = dict(os.environ, ZK_SPLIT_ATTN=variant)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        assert r.returncode == 0, (variant, r.stdout[-500:], r.stderr[-1500:])
