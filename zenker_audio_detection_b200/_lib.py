"""ctypes binding of ``libzk_b200.so`` (the C ABI declared in ``include/zk_b200.h``).

There is deliberately no fallback: if the shared library is missing or a call fails, a
``ZkError`` is raised.  PyTorch is used only for device memory and streams; every pointer
handed to the library is ``tensor.data_ptr()`` and every launch goes to
``torch.cuda.current_stream()``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

LIB_NAME = "libzk_b200.so"
# ZK_B200_LIB points at another build of the same library (same-box A/B runs of a kernel change); default: in-tree
LIB_PATH = os.environ.get("ZK_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", LIB_NAME)
AST_LAYERS = 12

EPI_BIAS_BF16, EPI_BIAS_GELU_BF16, EPI_BIAS_RESID_F32, EPI_PATCH_F32, EPI_BIAS_SPLIT, EPI_BIAS_GELU_SPLIT = 0, 1, 2, 3, 4, 5
FMT_BF16, FMT_F16 = 0, 1                  # zk_operand_format
PRECISION_FAST, PRECISION_RECHECK = 0, 1  # zk_precision
ABI_VERSION = 2


class ZkError(RuntimeError):
    pass


_fp = C.POINTER(C.c_float)


class AstLayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "ln1_w", "ln1_b", "q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "o_w", "o_b",
        "ln2_w", "ln2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b")]


class AstWeights(C.Structure):
    _fields_ = [
        ("num_layers", C.c_int32), ("max_length", C.c_int32), ("num_labels", C.c_int32), ("ln_eps", C.c_float),
        ("operand_format", C.c_int32),
        ("cls_token", C.c_void_p), ("dist_token", C.c_void_p), ("pos_emb", C.c_void_p),
        ("patch_w", C.c_void_p), ("patch_b", C.c_void_p),
        ("layer", AstLayerWeights * AST_LAYERS),
        ("final_ln_w", C.c_void_p), ("final_ln_b", C.c_void_p),
        ("head_ln_w", C.c_void_p), ("head_ln_b", C.c_void_p),
        ("head_w", C.c_void_p), ("head_b", C.c_void_p),
    ]


class CascadeParams(C.Structure):
    _fields_ = [("batch_size", C.c_int32), ("recheck_batch", C.c_int32), ("window_samples", C.c_int32),
                ("hop_samples", C.c_int32), ("mean1", C.c_float), ("std1", C.c_float), ("mean2", C.c_float),
                ("std2", C.c_float), ("thr1", C.c_float), ("min_prob", C.c_float), ("thr2", C.c_float),
                ("stage2_argmax", C.c_int32), ("recheck_eps", C.c_float)]


class CascadeCounts(C.Structure):
    _fields_ = [("num_windows", C.c_int32), ("num_forwarded", C.c_int32), ("rechecked_s1", C.c_int32),
                ("rechecked_s2", C.c_int32)]


# name -> (restype, argtypes); kept in one table so tests can check every symbol of the header is exported
SIGNATURES = {
    "zk_abi_version": (C.c_int, []),
    "zk_last_error_string": (C.c_char_p, []),
    "zk_device_check": (C.c_int, []),
    "zk_resample_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p, C.c_int64, C.c_void_p]),
    "zk_resample_pcm16": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_int64, C.c_void_p]),
    "zk_fbank_plan_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.POINTER(C.c_void_p)]),
    "zk_fbank_plan_destroy": (None, [C.c_void_p]),
    "zk_fbank_num_frames": (C.c_int64, [C.c_int64]),
    "zk_fbank_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "zk_fx_contract_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_float,
                                     C.c_float, C.c_int, C.c_void_p, C.c_void_p]),
    "zk_fx_stats_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "zk_model_create": (C.c_int, [C.POINTER(AstWeights), C.POINTER(C.c_void_p)]),
    "zk_model_destroy": (None, [C.c_void_p]),
    "zk_model_num_tokens": (C.c_int, [C.c_void_p]),
    "zk_model_max_length": (C.c_int, [C.c_void_p]),
    "zk_model_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "zk_model_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "zk_model_forward_fbank": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                         C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                                         C.c_void_p]),
    "zk_gate_compact": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "zk_softmax2": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "zk_band_select": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p]),
    "zk_scatter_rows2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "zk_cascade_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(CascadeParams)]),
    "zk_cascade_run": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(CascadeParams),
                                 C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.POINTER(CascadeCounts), C.c_void_p]),
    "zk_sum_sumsq_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "zk_gemm16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                            C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_void_p]),
    "zk_layernorm16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                 C.c_int, C.c_void_p]),
    "zk_attention16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "zk_attention_split": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "zk_f32_to_16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "zk_gemm_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                               C.c_void_p, C.c_int, C.c_void_p]),
    "zk_layernorm_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int64, C.c_int,
                                    C.c_void_p]),
    "zk_attention_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "zk_f32_to_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "zk_attention_trace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "zk_prof_enable": (None, [C.c_int]),
    "zk_prof_collect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "zk_kernel_class_name": (C.c_char_p, [C.c_int]),
}
NUM_KERNEL_CLASSES = 15


def prof_enable(time_launches: bool) -> None:
    load().zk_prof_enable(1 if time_launches else 0)


def prof_collect():
    """-> {class name: (ms, launches)}; synchronises the device and resets the counters."""
    lib = load()
    ms = (C.c_float * NUM_KERNEL_CLASSES)()
    n = (C.c_int64 * NUM_KERNEL_CLASSES)()
    check(lib.zk_prof_collect(ms, n), "zk_prof_collect")
    return {lib.zk_kernel_class_name(i).decode(): (float(ms[i]), int(n[i])) for i in range(NUM_KERNEL_CLASSES)}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load (once) and return the library; raises ``ZkError`` when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ZkError(
            f"{LIB_PATH} not found: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "There is no CPU or PyTorch fallback for the zenker-b200 path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.zk_abi_version() != ABI_VERSION:
        raise ZkError(f"ABI version mismatch: library {lib.zk_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def last_error() -> str:
    return load().zk_last_error_string().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise ZkError(f"{what} failed with status {rc}: {last_error()}")


def stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def require_device() -> None:
    """Raise unless the current CUDA device can run the sm_100a kernels."""
    import torch

    if not torch.cuda.is_available():
        raise ZkError("zenker-b200 needs a CUDA device (sm_100 / B200); there is no CPU path")
    check(load().zk_device_check(), "zk_device_check")
