"""Pins the numpy oracle to the reference: oracle(inputs) == golden outputs, where the goldens were
produced by running pypose v0.9.5 itself (oracle/make_golden.py).  Bit-level agreement is not
expected (different op order), 1e-13 absolute in fp64 is."""
import numpy as np
import pytest

from oracle import lie_oracle
from tests.util import all_ops, gold_case

KEYS = [k for k, *_ in all_ops()] + ["so3_jr"]


@pytest.mark.parametrize("key", KEYS)
def test_oracle_matches_reference_golden(golden, key):
    ins, outs = gold_case(golden, key)
    res = lie_oracle.run(key, *ins)
    assert len(res) == len(outs)
    for r, o in zip(res, outs):
        assert r.shape == o.shape
        np.testing.assert_allclose(r, o, rtol=0, atol=1e-13)


def test_golden_covers_every_abi_op(golden):
    have = {k.rsplit("/", 1)[0] for k in golden.files}
    assert set(KEYS) <= have


def test_pm_known_answers():
    # reference tests/basics/test_ops.py:9-13
    np.testing.assert_array_equal(lie_oracle.pm(np.array([0.1, 0.0, -0.2])), [1.0, 1.0, -1.0])


def test_config1_so3_roundtrip_fp64():
    """BASELINE.json configs[0]: SO3 Exp->Log round trip, batch 1024, fp64, CPU: <= 1e-12."""
    rng = np.random.default_rng(0)
    d = rng.standard_normal((1024, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    x = d * rng.uniform(0, np.pi - 1e-3, (1024, 1))
    x[:5] = d[:5] * np.array([[0.0], [1e-12], [1e-8], [1e-4], [np.pi - 1e-6]])
    rt = lie_oracle.log("SO3", lie_oracle.exp("SO3", x))
    assert np.abs(rt - x).max() <= 1e-12


def test_torch_port_matches_oracle():
    """oracle/torch_port.py (the multi-threaded CPU-baseline port) agrees with the numpy oracle."""
    import torch
    from oracle import torch_port
    from tests.util import rand_algebra
    rng = np.random.default_rng(3)
    x = rand_algebra(rng, "SE3", 2048)
    X = torch_port.se3_exp(torch.from_numpy(x))
    np.testing.assert_allclose(X.numpy(), lie_oracle.exp("SE3", x), atol=1e-13)
    np.testing.assert_allclose(torch_port.SE3_log(X).numpy(), lie_oracle.log("SE3", X.numpy()), atol=1e-13)
