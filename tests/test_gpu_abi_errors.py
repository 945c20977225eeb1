"""Error behaviour of the C ABI (include/zk_b200.h, "Errors"): bad arguments come back as a negative zk_status with a
message in zk_last_error_string(), nothing aborts or throws across the boundary, and the library stays usable for the
next, well-formed call.  Called through ctypes exactly as a foreign host would (no wrapper conveniences)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ERR_ARG, ERR_SHAPE, ERR_UNSUPPORTED, ERR_WORKSPACE = -1, -2, -3, -4


@pytest.fixture(scope="module")
def lib():
    from zenker_audio_detection_b200 import _lib

    _lib.require_device()
    return _lib.load()


def _msg(lib):
    return lib.zk_last_error_string().decode()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def test_version_and_device(lib):
    from zenker_audio_detection_b200 import _lib

    assert lib.zk_abi_version() == _lib.ABI_VERSION
    assert lib.zk_device_check() == 0
    assert lib.zk_kernel_class_name(6).decode() == "attention"
    assert lib.zk_fbank_num_frames(399) == 0 and lib.zk_fbank_num_frames(400) == 1 and lib.zk_fbank_num_frames(16000) == 98


def test_attention_rejects_bad_arguments_and_recovers(lib):
    T = 130
    qkv = torch.randn(T, 2304, device="cuda").half()
    out = torch.empty(T, 768, device="cuda", dtype=torch.float16)
    assert lib.zk_attention16(None, out.data_ptr(), 1, T, 1, _stream()) == ERR_ARG and "null" in _msg(lib)
    assert lib.zk_attention16(qkv.data_ptr(), out.data_ptr(), 0, T, 1, _stream()) == ERR_ARG
    assert lib.zk_attention16(qkv.data_ptr(), out.data_ptr(), 1, T, 7, _stream()) == ERR_ARG and "format" in _msg(lib)
    assert lib.zk_attention16(qkv.data_ptr(), out.data_ptr(), 1 << 30, 1 << 12, 1, _stream()) == ERR_SHAPE
    # a well-formed call right after: same result as before the failures
    assert lib.zk_attention16(qkv.data_ptr(), out.data_ptr(), 1, T, 1, _stream()) == 0
    q, k, v = (qkv[:, i * 768:(i + 1) * 768].float().view(1, T, 12, 64).transpose(1, 2) for i in range(3))
    ref = (torch.softmax((q @ k.transpose(2, 3)) * 0.125, -1) @ v).transpose(1, 2).reshape(T, 768)
    assert (out.float() - ref).abs().max().item() < 5e-3


def test_split_attention_alignment(lib):
    T = 64
    buf = torch.zeros(T * 4608 + 8, device="cuda", dtype=torch.float16)
    out = torch.zeros(T * 1536 + 8, device="cuda", dtype=torch.float16)
    assert lib.zk_attention_split(buf.data_ptr() + 2, out.data_ptr(), 1, T, _stream()) == ERR_ARG and "aligned" in _msg(lib)
    assert lib.zk_attention_split(buf.data_ptr(), out.data_ptr() + 2, 1, T, _stream()) == ERR_ARG
    assert lib.zk_attention_split(buf.data_ptr(), out.data_ptr(), 1, T, _stream()) == 0
    torch.cuda.synchronize()


def test_layernorm_and_convert_argument_checks(lib):
    x = torch.randn(8, 768, device="cuda")
    w = torch.ones(768, device="cuda")
    o = torch.empty(8, 1536, device="cuda", dtype=torch.float16)
    assert lib.zk_layernorm16(x.data_ptr(), w.data_ptr(), w.data_ptr(), 1e-12, o.data_ptr(), 8, 512, 1, 1, _stream()) == ERR_SHAPE
    assert "768" in _msg(lib)
    assert lib.zk_layernorm16(x.data_ptr(), w.data_ptr(), w.data_ptr(), 1e-12, o.data_ptr(), 8, 768, 0, 2, _stream()) == ERR_ARG  # bf16 x 2 planes
    assert lib.zk_layernorm16(x.data_ptr(), None, w.data_ptr(), 1e-12, o.data_ptr(), 8, 768, 1, 1, _stream()) == ERR_ARG
    assert lib.zk_layernorm16(x.data_ptr(), w.data_ptr(), w.data_ptr(), 1e-12, o.data_ptr(), 8, 768, 1, 2, _stream()) == 0
    assert lib.zk_f32_to_16(x.data_ptr(), o.data_ptr(), 8, 770, 1, 1, 1.0, _stream()) == ERR_ARG and "multiple of 4" in _msg(lib)
    assert lib.zk_f32_to_16(x.data_ptr(), o.data_ptr(), 8, 768, 0, 2, 1.0, _stream()) == ERR_ARG      # two planes are fp16 only
    assert lib.zk_f32_to_16(x.data_ptr(), o.data_ptr(), 0, 768, 1, 1, 1.0, _stream()) == 0            # nothing to do
    assert lib.zk_f32_to_16(x.data_ptr(), o.data_ptr(), 8, 768, 1, 2, 1.0, _stream()) == 0
    torch.cuda.synchronize()
    hi, lo = o[:, :768].float(), o[:, 768:].float()
    assert (hi + lo - x).abs().max().item() < 1e-6 * x.abs().max().item() + 1e-7


def test_gemm_shape_and_pointer_checks(lib):
    from zenker_audio_detection_b200 import _lib

    M, N, K = 256, 256, 128
    a = torch.randn(M, K, device="cuda").half()
    w = torch.randn(N, K, device="cuda").half()
    b = torch.zeros(N, device="cuda")
    o = torch.empty(M, N, device="cuda", dtype=torch.float16)

    def call(a_ptr, M_, N_, K_, epi=_lib.EPI_BIAS_BF16, fmt=1, products=1):
        return lib.zk_gemm16(a_ptr, 0, w.data_ptr(), 0, b.data_ptr(), o.data_ptr(), 0, M_, N_, K_, epi, fmt, products, 1.0,
                             None, 0, _stream())

    assert call(None, M, N, K) == ERR_ARG
    assert call(a.data_ptr(), M, N, 100) in (ERR_SHAPE, ERR_ARG) and "K" in _msg(lib)     # K must be a multiple of 64
    assert call(a.data_ptr(), M, 100, K) in (ERR_SHAPE, ERR_ARG)
    assert call(a.data_ptr(), M, N, K, epi=99) in (ERR_ARG, ERR_SHAPE)
    assert call(a.data_ptr(), M, N, K, products=2) in (ERR_ARG, ERR_SHAPE)
    assert call(a.data_ptr(), M, N, K) == 0
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    assert (o.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()


def test_fbank_plan_and_frontend_checks(lib):
    from zenker_audio_detection_b200 import ops

    plan = ops.FbankPlan()
    wave = torch.randn(16000, device="cuda")
    m = int(lib.zk_fbank_num_frames(16000))
    out = torch.empty(m, 128, device="cuda")
    assert lib.zk_fbank_f32(None, wave.data_ptr(), 16000, out.data_ptr(), m, _stream()) == ERR_ARG
    assert lib.zk_fbank_f32(plan._h, wave.data_ptr(), 16000, out.data_ptr(), m + 5, _stream()) in (ERR_ARG, ERR_SHAPE)
    assert lib.zk_fbank_f32(plan._h, wave.data_ptr(), 16000, out.data_ptr(), m, _stream()) == 0
    # a plan belongs to the device it was created on: using it elsewhere is an argument error, not a fault
    if torch.cuda.device_count() > 1:
        with torch.cuda.device(1):
            w1 = torch.randn(16000, device="cuda:1")
            o1 = torch.empty(m, 128, device="cuda:1")
            rc = lib.zk_fbank_f32(plan._h, w1.data_ptr(), 16000, o1.data_ptr(), m, torch.cuda.current_stream().cuda_stream)
            assert rc == ERR_ARG and "device" in _msg(lib)
    # resampler: taps are required, ratios must be positive
    o16 = torch.empty(5334, device="cuda")
    assert lib.zk_resample_f32(wave.data_ptr(), 16000, 1, 16000, None, 3, 1, 19, o16.data_ptr(), 5334, _stream()) == ERR_ARG
    torch.cuda.synchronize()


def test_model_workspace_and_precision_checks(lib):
    from zenker_audio_detection_b200 import _lib, ops, synth

    m = ops.AstModel(synth.random_state_dict(3), num_layers=1)
    feats = torch.zeros(2, 1024, 128, device="cuda")
    logits = torch.empty(2, 2, device="cuda")
    need = lib.zk_model_workspace_bytes(m._h, 2, _lib.PRECISION_FAST)
    assert need > 0 and lib.zk_model_workspace_bytes(m._h, 2, _lib.PRECISION_RECHECK) > need
    assert lib.zk_model_workspace_bytes(None, 2, 0) == 0
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    args = (m._h, feats.data_ptr(), None, 2)
    assert lib.zk_model_forward(*args, 5, ws.data_ptr(), need, logits.data_ptr(), None, _stream()) == ERR_ARG and "precision" in _msg(lib)
    assert lib.zk_model_forward(*args, 0, ws.data_ptr(), need // 2, logits.data_ptr(), None, _stream()) == ERR_WORKSPACE
    assert "workspace" in _msg(lib)
    assert lib.zk_model_forward(*args, 0, ws.data_ptr() + 8, need - 8, logits.data_ptr(), None, _stream()) == ERR_ARG  # 256-B alignment
    assert lib.zk_model_forward(*args, 0, ws.data_ptr(), need, logits.data_ptr(), None, _stream()) == 0
    torch.cuda.synchronize()
    assert torch.isfinite(logits).all()


def test_gate_and_band_select_checks(lib):
    n = 10
    logits = torch.randn(n, 2, device="cuda")
    probs = torch.empty(n, 2, device="cuda")
    pred = torch.empty(n, dtype=torch.int32, device="cuda")
    idx = torch.empty(n, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    assert lib.zk_gate_compact(None, n, 0.5, -1.0, probs.data_ptr(), pred.data_ptr(), idx.data_ptr(), cnt.data_ptr(), _stream()) == ERR_ARG
    assert lib.zk_gate_compact(logits.data_ptr(), -1, 0.5, -1.0, probs.data_ptr(), pred.data_ptr(), idx.data_ptr(), cnt.data_ptr(), _stream()) == ERR_ARG
    assert lib.zk_gate_compact(logits.data_ptr(), n, 0.5, -1.0, probs.data_ptr(), pred.data_ptr(), idx.data_ptr(), cnt.data_ptr(), _stream()) == 0
    torch.cuda.synchronize()
    p = torch.softmax(logits, 1)
    want = ((p[:, 1] > p[:, 0]) & (p[:, 1] >= 0.5)).nonzero().flatten().cpu().numpy()
    assert int(cnt.item()) == len(want) and np.array_equal(idx[:len(want)].cpu().numpy(), want)
    margins = (C.c_float * 5)(0.0, 0.1, 0.2, 0.3, 0.4)
    pos = torch.empty(n, dtype=torch.int32, device="cuda")
    win = torch.empty(n, dtype=torch.int32, device="cuda")
    rc = lib.zk_band_select(logits.data_ptr(), n, margins, 5, 0.01, None, pos.data_ptr(), win.data_ptr(), cnt.data_ptr(), _stream())
    assert rc in (ERR_ARG, ERR_SHAPE) and "4" in _msg(lib)      # at most four decision points per stage


def test_plain_c_host_runs_the_cascade(c_host_binary):
    """examples/cascade_host.c on the GPU: a C99 program with no Python and no torch in the process resamples a stereo
    PCM16 recording and runs the whole two-stage cascade through zk_resample_pcm16 + zk_cascade_run, and its own checks
    (window count of ref:62-75, probabilities summing to one, gate mask and ascending index list as functions of the
    returned probabilities, ref:312-320) hold: exit code 0."""
    import re
    import subprocess

    r = subprocess.run([c_host_binary, "12"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    m = re.search(r"-> (\d+) windows, (\d+) forwarded to stage 2 .* checks ok", r.stdout)
    assert m, r.stdout
    assert int(m.group(1)) == 23 and 0 <= int(m.group(2)) <= 23      # 12 s at hop 0.5 s -> 23 windows
    print(r.stdout.strip())
