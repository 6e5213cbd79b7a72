"""CUDA row_sparse push / row_sparse_pull against outputs of the REFERENCE's own code
(tests/golden/rowsparse_reduce_retain.npz, written by oracle/gen_golden.py through
oracle/ref_sparse.cc: ElementwiseSumRsp, src/ndarray/ndarray_function.cc:59-176; the sparse_retain
kernels, src/operator/tensor/sparse_retain-inl.h:121-262; UniqueImpl, src/kvstore/kvstore_utils.cc:
31-44). No oracle in between: ids and rows bit for bit. The store applies Unique to the row ids
before the retain (kvstore_local.h:415-446), so a pull of unsorted / repeated ids is compared with
the fixture's rows at the first occurrence of each id."""
import numpy as np
import pytest

from gen_golden import rsp_cases

pytestmark = pytest.mark.gpu

ROWS, RL = 211, 19


@pytest.fixture(scope="module")
def mx():
    import anand_mxnet_b200 as mx
    return mx


def eq(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def rsp(mx, idx, val):
    if len(idx) == 0:
        return mx.nd.sparse.zeros('row_sparse', (ROWS, RL), mx.gpu(0))
    return mx.nd.sparse.row_sparse_array((val, idx), shape=(ROWS, RL), ctx=mx.gpu(0))


def pull(mx, kv, key, ids):
    out = mx.nd.sparse.zeros('row_sparse', (ROWS, RL), mx.gpu(0))
    kv.row_sparse_pull(key, out=out, row_ids=mx.nd.array(np.ascontiguousarray(ids, np.int64), mx.gpu(0), np.int64))
    return out.indices.asnumpy(), out.data.asnumpy()


@pytest.mark.parametrize("tag", ["dup4", "empty_mid", "full_plus", "single", "nine", "all_empty"])
def test_push_merge_golden(mx, golden, tag):
    g = golden("rowsparse_reduce_retain")
    idxs, vals = rsp_cases()[0][tag]
    want_idx, want_val = g["reduce_%s_idx" % tag], g["reduce_%s_val" % tag]
    kv = mx.kv.create('device')
    kv.init(tag, mx.nd.sparse.zeros('row_sparse', (ROWS, RL), mx.gpu(0)))
    kv.push(tag, [rsp(mx, i, v) for i, v in zip(idxs, vals)])       # no optimizer: stored = merged
    got_idx, got_val = pull(mx, kv, tag, np.arange(ROWS))
    if len(want_idx) == 0:
        assert got_idx.shape[0] == 0                                  # FillZerosRspImpl: nothing stored
        return
    dense = np.zeros((ROWS, RL), np.float32)
    dense[want_idx] = want_val
    assert np.array_equal(got_idx, np.arange(ROWS))
    assert eq(got_val.reshape(ROWS, RL), dense)
    # only the merged rows, asked for in descending order
    got_idx, got_val = pull(mx, kv, tag, want_idx[::-1])
    assert np.array_equal(got_idx, want_idx)
    assert eq(got_val.reshape(len(want_idx), RL), want_val)


@pytest.mark.parametrize("tag", ["sorted_unique", "unsorted_dup", "none_present", "dense_src"])
def test_pull_retain_golden(mx, golden, tag):
    g = golden("rowsparse_reduce_retain")
    src_i, src_v, ids, dense_src = rsp_cases()[1][tag]
    want_idx, want_val = g["retain_%s_idx" % tag], g["retain_%s_val" % tag]
    assert np.array_equal(want_idx, ids)                 # the kernel emits every requested id
    kv = mx.kv.create('device')
    if dense_src:
        kv.init(tag, mx.nd.array(src_v, mx.gpu(0)).tostype('row_sparse'))
    else:
        kv.init(tag, mx.nd.sparse.zeros('row_sparse', (ROWS, RL), mx.gpu(0)))
        kv.push(tag, [rsp(mx, src_i, src_v)])
    uniq, first = np.unique(ids, return_index=True)
    got_idx, got_val = pull(mx, kv, tag, ids)
    assert np.array_equal(got_idx, uniq)
    assert eq(got_val.reshape(len(uniq), RL), want_val[first])
