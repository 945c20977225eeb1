"""tcgen05 GEMM + fused epilogues vs torch on the same 16-bit-rounded operands (through the C ABI): bf16 and fp16
single-product GEMMs against fp32, the split-operand (hi | lo fp16 planes, three products) GEMM against float64."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(M, N, K, seed, dtype=torch.bfloat16):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(dtype)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(dtype)
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    return a, w, b


def _ref(a, w, b):
    return a.float() @ w.float().t() + b


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (2465, 768, 768), (1214 * 3, 2304, 768), (300, 3072, 768),
                                   (1214 * 2 + 5, 768, 3072), (1, 256, 128)])
def test_gemm_bias_bf16(M, N, K):
    from zenker_audio_detection_b200 import _lib, ops

    a, w, b = _mk(M, N, K, 1)
    out = ops.gemm(a, w, b, _lib.EPI_BIAS_BF16)
    torch.cuda.synchronize()
    ref = _ref(a, w, b)
    err = (out.float() - ref).abs().max().item()
    tol = 8e-3 * ref.abs().max().item() + 1e-3  # bf16 output rounding (2^-8 relative)
    assert err <= tol, (err, tol)


def test_gemm_gelu():
    from zenker_audio_detection_b200 import _lib, ops

    a, w, b = _mk(1214 + 77, 3072, 768, 2)
    out = ops.gemm(a, w, b, _lib.EPI_BIAS_GELU_BF16)
    ref = torch.nn.functional.gelu(_ref(a, w, b))
    err = (out.float() - ref).abs().max().item()
    assert err <= 8e-3 * ref.abs().max().item() + 1e-3, err


def test_gemm_residual_f32():
    from zenker_audio_detection_b200 import _lib, ops

    a, w, b = _mk(2000, 768, 3072, 3)
    x = torch.randn(2000, 768, device="cuda")
    x0 = x.clone()
    ops.gemm(a, w, b, _lib.EPI_BIAS_RESID_F32, out=x)
    ref = x0 + _ref(a, w, b)
    err = (x - ref).abs().max().item()
    assert err <= 2e-4 * ref.abs().max().item() + 1e-4, err  # fp32 accumulate, order differs only


def test_gemm_patch_scatter():
    from zenker_audio_detection_b200 import _lib, ops

    B, P, T = 3, 1212, 1214
    a, w, b = _mk(B * P, 768, 256, 4)
    pos = torch.randn(T, 768, device="cuda")
    x = torch.full((B * T, 768), 7.0, device="cuda")
    ops.gemm(a, w, b, _lib.EPI_PATCH_F32, out=x, aux=pos, aux_rows=P)
    ref = (_ref(a, w, b).view(B, P, 768) + pos[2:].unsqueeze(0))
    got = x.view(B, T, 768)
    assert torch.all(got[:, :2] == 7.0)  # cls/dist rows untouched
    err = (got[:, 2:] - ref).abs().max().item()
    assert err <= 2e-4 * ref.abs().max().item() + 1e-4, err


def test_gemm_rejects_bad_shapes():
    from zenker_audio_detection_b200 import _lib, ops

    a, w, b = _mk(64, 100, 64, 5)
    with pytest.raises(_lib.ZkError):
        ops.gemm(a, w, b, _lib.EPI_BIAS_BF16)


# ------------------------------------------------------------------------------------------------ fp16 operands
@pytest.mark.parametrize("M,N,K", [(2465, 768, 768), (1214 * 3, 2304, 768), (1214 * 2 + 5, 768, 3072), (1, 256, 128)])
def test_gemm_bias_fp16(M, N, K):
    from zenker_audio_detection_b200 import _lib, ops

    a, w, b = _mk(M, N, K, 11, torch.float16)
    out = ops.gemm(a, w, b, _lib.EPI_BIAS_BF16, acc_scale=0.5)
    assert out.dtype == torch.float16
    ref = 0.5 * (a.float() @ w.float().t()) + b
    err = (out.float() - ref).abs().max().item()
    tol = 1e-3 * ref.abs().max().item() + 1e-4  # fp16 output rounding (2^-11 relative)
    assert err <= tol, (err, tol)


def test_gemm_gelu_fp16():
    from zenker_audio_detection_b200 import _lib, ops

    a, w, b = _mk(1214 + 77, 3072, 768, 12, torch.float16)
    out = ops.gemm(a, w, b, _lib.EPI_BIAS_GELU_BF16)
    ref = torch.nn.functional.gelu(_ref(a, w, b))
    err = (out.float() - ref).abs().max().item()
    assert err <= 1e-3 * ref.abs().max().item() + 1e-4, err


def test_gemm_fp16_saturates_instead_of_inf():
    from zenker_audio_detection_b200 import _lib, ops

    a = torch.full((128, 64), 100.0, device="cuda", dtype=torch.float16)
    w = torch.full((256, 64), 100.0, device="cuda", dtype=torch.float16)
    out = ops.gemm(a, w, torch.zeros(256, device="cuda"), _lib.EPI_BIAS_BF16)  # 640 000 > 65 504
    assert torch.isfinite(out.float()).all() and float(out.float().max()) == 65504.0


# ------------------------------------------------------------------------------------------------ split operands
def _split_ref(x64, scale=1.0):
    """float64 value of the hi + lo fp16 planes zk_f32_to_16 writes (what the GEMM really multiplies)."""
    x = (x64 * scale).float()
    hi = x.half()
    lo = (x - hi.float()).half()
    return (hi.double() + lo.double()) / scale


@pytest.mark.parametrize("M,N,K,wstd", [(1214 * 2 + 5, 768, 768, 0.02), (700, 2304, 768, 0.08), (515, 768, 3072, 0.02),
                                        (130, 768, 256, 0.02), (3, 256, 64, 1.0)])
def test_split_gemm_matches_float64(M, N, K, wstd):
    """x2 mode: A, W as fp16 hi | lo planes, C = A_lo W_hi + A_hi W_lo + A_hi W_hi.  Against float64 on the fp32 inputs
    the error must be fp32-class: dropping the lo*lo term and the plane rounding are each ~2^-22 relative per product."""
    from zenker_audio_detection_b200 import _lib, ops

    g = torch.Generator(device="cuda").manual_seed(M + N)
    a32 = torch.randn(M, K, device="cuda", generator=g)
    w32 = torch.randn(N, K, device="cuda", generator=g) * wstd
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    scale = 2.0 ** torch.floor(torch.log2(16384.0 / w32.abs().max())).item()
    a2, w2 = ops.split_f16(a32), ops.split_f16(w32, scale)
    x = torch.randn(M, N, device="cuda", generator=g)
    x0 = x.clone()
    ops.gemm(a2, w2, b, _lib.EPI_BIAS_RESID_F32, out=x, products=3, acc_scale=1.0 / scale)
    ref = x0.double() + a32.double() @ w32.double().t() + b.double()
    mag = (a32.double().abs() @ w32.double().abs().t()).max().item()  # sum |a||w|: what rounding errors scale with
    err = (x.double() - ref).abs().max().item()
    assert err <= 3e-6 * mag, (err, mag)
    # and the same contraction through plain fp32 torch is not better than a few times this
    err32 = ((x0 + a32 @ w32.t() + b).double() - ref).abs().max().item()
    rms = (x.double() - ref).pow(2).mean().sqrt().item()
    rms32 = ((x0 + a32 @ w32.t() + b).double() - ref).pow(2).mean().sqrt().item()
    print(f"split gemm {M}x{N}x{K}: max err {err:.3e} (torch fp32 {err32:.3e}), rms {rms:.3e} (torch fp32 {rms32:.3e}), "
          f"sum|a||w| {mag:.3e}")
    # the segmented chains (K = 768, 3072) must be as good as an fp32 FFMA matmul, not just "fp32-class"
    if K % 256 == 0 and K > 256:
        assert rms <= 1.5 * rms32 + 1e-9, (rms, rms32)
    # bit-reproducible: the per-segment reduce-adds of a tile always land in the same order
    y = x0.clone()
    ops.gemm(a2, w2, b, _lib.EPI_BIAS_RESID_F32, out=y, products=3, acc_scale=1.0 / scale)
    assert torch.equal(x, y)


def test_split_gemm_split_output_and_exact_gelu():
    from zenker_audio_detection_b200 import _lib, ops

    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 1214 + 9, 3072, 768
    a32 = torch.randn(M, K, device="cuda", generator=g)
    w32 = torch.randn(N, K, device="cuda", generator=g) * 0.03
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    scale = 2.0 ** torch.floor(torch.log2(16384.0 / w32.abs().max())).item()
    a2, w2 = ops.split_f16(a32), ops.split_f16(w32, scale)
    pre = a32.double() @ w32.double().t() + b.double()
    for epi, ref in ((_lib.EPI_BIAS_SPLIT, pre), (_lib.EPI_BIAS_GELU_SPLIT, torch.nn.functional.gelu(pre))):
        out = ops.gemm(a2, w2, b, epi, products=3, acc_scale=1.0 / scale)
        assert out.shape == (M, 2 * N) and out.dtype == torch.float16
        val = out[:, :N].double() + out[:, N:].double()
        err = (val - ref).abs().max().item()
        assert err <= 1e-5 * ref.abs().max().item() + 1e-6, (epi, err)
        # the hi plane alone is the fp16 rounding of the value
        assert (out[:, :N].float() - val.float()).abs().max().item() <= 1e-3 * ref.abs().max().item()


def test_split_f16_planes():
    from zenker_audio_detection_b200 import ops

    x = torch.randn(77, 768, device="cuda") * torch.logspace(-6, 2, 768, device="cuda")
    p = ops.split_f16(x, 4.0)
    hi, lo = p[:, :768], p[:, 768:]
    assert torch.equal(hi, (x * 4).half())
    assert torch.equal(lo, (x * 4 - hi.float()).half())
    val = (hi.double() + lo.double()) / 4
    assert ((val - x.double()).abs() <= 2.0 ** -22 * x.double().abs() + 2.0 ** -26).all()
