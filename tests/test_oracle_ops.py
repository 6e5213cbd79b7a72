"""Pins the multi-tensor optimizer restatements of oracle/kvoracle.c (LARS, AdamW,
LAMB) bit-for-bit against (a) the reference's own FCompute<cpu> functions compiled in place
(oracle/_ref/libmxref.so, oracle/ref_ops.cc) when present, and (b) the committed fixture
tests/golden/multi_tensor_ops.npz generated from them by oracle/gen_golden.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import kvoracle as K  # noqa: E402
import golden_ops as G  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "multi_tensor_ops.npz")


def _same(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and \
        np.array_equal(a.view(np.uint8), b.view(np.uint8))


@pytest.mark.parametrize("case", sorted(G.CASES))
def test_oracle_matches_golden(case):
    gold = np.load(GOLDEN)
    got = G.run_case(case, K.get_oracle())
    names = [k for k in gold.files if k.startswith(case + "/")]
    assert names, "no golden entries for %s" % case
    assert sorted(names) == sorted(case + "/" + k for k in got)
    for k, v in got.items():
        assert _same(v, gold[case + "/" + k]), (case, k)


@pytest.mark.parametrize("case", sorted(G.CASES))
def test_oracle_matches_live_reference(case):
    ref = K.ref()
    if ref is None or not ref.has_ops():
        pytest.skip("oracle/_ref/libmxref.so not built here")
    want = G.run_case(case, None, ref=ref)
    got = G.run_case(case, K.get_oracle())
    assert sorted(want) == sorted(got)
    for k in want:
        assert _same(got[k], want[k]), (case, k)
