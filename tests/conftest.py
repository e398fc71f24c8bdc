import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # A fresh checkout has no built artefacts (they are git-ignored): build the CUDA library, the C oracle and --
    # when /root/reference is present -- the reference's own kernel once, so that the suite is self-sufficient.
    lib = os.path.join(ROOT, "nd_b200", "libndnlm.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected explicitly with `-m gpu`; without a device they are skipped, never faked.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import json
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "nlm_golden.npz")
    z = np.load(path)
    meta = json.loads(bytes(z["__meta__"]).decode())
    return z, meta


@pytest.fixture(scope="session")
def c_oracle():
    from oracle import c_port
    c_port.build()
    return c_port
