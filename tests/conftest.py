import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `pytest -m gpu` under gpurun)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def c_host_binary(tmp_path_factory):
    """examples/cascade_host.c built with plain gcc (C99, warnings are errors) against include/zk_b200.h and the in-tree
    libzk_b200.so: the proof that the boundary is consumable without Python or torch.  Built with gcc directly, not
    through make, so that the library itself is never rebuilt on the GPU box."""
    import subprocess

    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    libdir = os.path.join(ROOT, "zenker_audio_detection_b200", "lib")
    if not os.path.exists(os.path.join(libdir, "libzk_b200.so")):
        import __graft_entry__ as g

        g.build()
    out = str(tmp_path_factory.mktemp("c_host") / "cascade_host")
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
           "-isystem", os.path.join(cuda, "include"), os.path.join(ROOT, "examples", "cascade_host.c"), "-o", out,
           "-L" + libdir, "-lzk_b200", "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-lm",
           "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.join(cuda, "lib64")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return out
