import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # The shared library is a build artefact (git-ignored): build it once if a fresh checkout has none (nvcc cross-compiles
    # sm_100a without a GPU).  Nothing falls back to another implementation when this fails — the tests then fail loudly.
    lib = os.path.join(ROOT, "sequoia_pub_b200", "libsequoia_b200.so")
    if not os.path.exists(lib):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "sequoia_pub_b200", "csrc"), "-j", str(os.cpu_count() or 4)], check=False)


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected explicitly with `-m gpu`; without a device they are skipped, never faked.
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
