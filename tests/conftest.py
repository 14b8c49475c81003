"""pytest configuration.

* registers the `gpu` marker (tests that need a real B200);
* gives the `b200pose::*` torch ops a **test-only CPU kernel backed by the oracle**, so the
  host-side logic (LieTensor dispatch, autograd/vmap registrations, LM control flow, gloo
  sharding) can be exercised without a GPU.  The shipped package registers CUDA kernels only —
  without this conftest a CPU tensor raises in the dispatcher (tests/test_abi.py checks that in a
  clean subprocess).
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _install_cpu_oracle_kernels():
    import pypose_b200  # noqa: F401  (defines the ops)
    from pypose_b200.lietensor import ops as _ops
    from oracle import lie_oracle

    def make(name):
        sym = name

        def cpu_impl(*ts):
            arrs = [t.detach().contiguous().numpy() for t in ts]
            outs = lie_oracle.run(sym, *arrs)
            outs = [torch.from_numpy(np.ascontiguousarray(o)).to(ts[0].dtype) for o in outs]
            return outs[0] if len(outs) == 1 else tuple(outs)
        return cpu_impl

    for name in _ops.OP_INFO:
        torch.library.impl(f"b200pose::{name}", "CPU")(make(name))
    try:
        from pypose_b200 import _testhooks
        _testhooks.install_cpu_oracle()
    except ImportError:
        pass


_install_cpu_oracle_kernels()


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "lie_ops.npz"))
