"""Committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle and
cross-checked there against the NumPy restatement): the oracle must reproduce them on any host (CPU test) and the
CUDA engine must match them at the north-star tolerance 1e-9 (GPU test, through the C ABI)."""
from __future__ import annotations

import importlib.util
import os

import numpy as np
import pytest

import parity as P
from oracle.oracle_lib import OracleBatch

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)

NAMES = sorted(G.SCENARIOS)


def load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(name):
    kind, B, script = G.SCENARIOS[name]
    fx = load(name)
    assert int(fx["kind"]) == kind and int(fx["B"]) == B
    o = G.make(OracleBatch, kind, B)
    script(o, B)
    # libm may differ in the last ulp between hosts; 1e-12 is far inside what any algorithmic change would cause
    P.assert_parity(kind, o.get_state(), (fx["mu"], fx["sigma"]), tol=1e-12, what=name)
    assert np.array_equal(o.get_status(), fx["status"]) and np.array_equal(o.get_last_time(), fx["last_time"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_engine_matches_golden(name):
    from slam_pose_estimation_b200 import UkfBatch

    kind, B, script = G.SCENARIOS[name]
    fx = load(name)
    g = G.make(UkfBatch, kind, B)
    script(g, B)
    P.assert_parity(kind, g.get_state(), (fx["mu"], fx["sigma"]), tol=P.TOL, what=name)
    assert np.array_equal(g.get_status(), fx["status"]) and np.array_equal(g.get_last_time(), fx["last_time"])
    assert g.launch_count() > 0
