"""tcgen05 GEMM + fused epilogues vs torch fp32 on the same bf16-rounded operands (through the C ABI)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(M, N, K, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
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
