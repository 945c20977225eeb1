"""The C-ABI library builds, loads on a GPU-less host and exports every symbol include/zk_b200.h declares
(no compute calls here), and the product path fails loudly -- never silently -- without a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    from zenker_audio_detection_b200 import _lib

    return _lib.load()


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "zk_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(zk_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from zenker_audio_detection_b200 import _lib

    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in zk_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(syms)


def test_abi_version_and_pure_host_entry_points(lib):
    assert lib.zk_abi_version() == 2
    assert lib.zk_fbank_num_frames(16000) == 98
    assert lib.zk_fbank_num_frames(399) == 0 and lib.zk_fbank_num_frames(400) == 1
    assert lib.zk_fbank_num_frames(9_600_000) == 59998


def test_struct_layout_matches_header(lib):
    from zenker_audio_detection_b200 import _lib

    assert ctypes.sizeof(_lib.AstLayerWeights) == 16 * 8
    assert ctypes.sizeof(_lib.AstWeights) == 16 + 8 + 5 * 8 + 12 * 16 * 8 + 6 * 8  # operand_format + padding
    assert _lib.AstWeights.cls_token.offset == 24


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_gpu(lib):
    from zenker_audio_detection_b200 import ZkError, ops, synth
    from zenker_audio_detection_b200.fx import ZenkerASTFeatureExtractor
    from zenker_audio_detection_b200.model import ZenkerASTForAudioClassification

    with pytest.raises(ZkError):
        ops.FbankPlan()
    with pytest.raises(ZkError):
        ops.resample(torch.zeros(100), 48000, 16000)
    with pytest.raises((ZkError, RuntimeError, AssertionError)):
        ZenkerASTFeatureExtractor()(synth.cfg1_windows(1)[0], sampling_rate=16000, return_tensors="pt")
    m = ZenkerASTForAudioClassification({"max_length": 1024}, synth.random_state_dict(0))
    with pytest.raises(ZkError):
        m.to("cpu")
    assert lib.zk_device_check() != 0 and b"" != lib.zk_last_error_string()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "zenker_audio_detection_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                # tables.py builds the constant tables with torch ops; compat.py imports torchaudio only to REBIND
                # torchaudio.load / torchaudio.info to the RIFF reader -- no torchaudio compute anywhere in the product
                assert "torchaudio" not in src or fn in ("tables.py", "compat.py") or "import torchaudio" not in src, fn


def test_a_plain_c_host_builds_against_the_header_and_fails_loudly_without_a_gpu(c_host_binary):
    """examples/cascade_host.c (C99, -Wall -Wextra -pedantic -Werror) compiles against include/zk_b200.h and links the
    in-tree library: the header is C, not C++ or torch.  Without an sm_100 device the program stops at zk_device_check
    with a message and a non-zero exit code -- it never computes anything on the CPU."""
    import subprocess

    if torch.cuda.is_available():
        pytest.skip("the GPU run of the example is tests/test_gpu_abi_errors.py::test_plain_c_host_runs_the_cascade")
    r = subprocess.run([c_host_binary, "2"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "zk_device_check" in r.stderr and "cascade_host:" not in r.stdout, (r.stdout, r.stderr)
