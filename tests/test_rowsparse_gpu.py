"""GPU parity tests of the row_sparse path (push of row_sparse gradients, row_sparse_pull, lazy
optimizer updates), through the C ABI. Modeled on tests/python/unittest/test_kvstore.py:69-97,178-227
and tests/python/gpu/test_kvstore_gpu.py:46-134 of the reference; checker = CPU oracle; ids and
pulled rows must be BIT-EXACT (north_star: "bit-exact for index/row_sparse pull")."""
import numpy as np
import pytest

import kvoracle as K

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mx():
    import anand_mxnet_b200 as mx
    return mx


def eq(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


@pytest.fixture(params=["bitmap", "sort"])
def union_path(request, monkeypatch):
    """both id-union front-ends of a push / pull: the bitmap rank (tables up to 8 M rows) and the
    radix sort of (id, position) pairs it falls back to (B200KV_RSP_SORT=1 forces it)"""
    if request.param == "sort":
        monkeypatch.setenv("B200KV_RSP_SORT", "1")
    else:
        monkeypatch.delenv("B200KV_RSP_SORT", raising=False)
    return request.param


def make_rsp(mx, rng, shape, nnr, ctx, integer=False):
    idx = np.sort(rng.choice(shape[0], nnr, replace=False)).astype(np.int64)
    if integer:
        val = rng.integers(-3, 4, (nnr,) + shape[1:]).astype(np.float32)
    else:
        val = rng.uniform(-1, 1, (nnr,) + shape[1:]).astype(np.float32)
    return mx.nd.sparse.row_sparse_array((val, idx), shape=shape, ctx=ctx), idx, val


def test_row_sparse_pull_from_dense_init(mx):
    # test_kvstore.py:69-97: init with a dense-valued rsp, pull random (duplicate, unsorted) ids
    rng = np.random.default_rng(0)
    shape = (64, 8)
    kv = mx.kv.create('local')
    w = rng.uniform(-1, 1, shape).astype(np.float32)
    kv.init('e', mx.nd.array(w, mx.gpu(0)).tostype('row_sparse'))
    for count in (1, 4):
        outs = [mx.nd.sparse.zeros('row_sparse', shape, mx.gpu(0)) for _ in range(count)]
        ids = [rng.integers(0, shape[0], 20) for _ in range(count)]
        # ids given as float32 arrays, as mx.nd.array() would create them in the reference's test
        kv.row_sparse_pull('e', out=outs, row_ids=[mx.nd.array(i.astype(np.float32), mx.gpu(0)) for i in ids])
        for o, i in zip(outs, ids):
            u = np.unique(i)
            assert np.array_equal(o.indices.asnumpy(), u)       # ascending unique, int64
            assert o.indices.asnumpy().dtype == np.int64
            assert eq(o.data.asnumpy(), w[u])
            dense = o.asnumpy()
            mask = np.zeros(shape[0], bool)
            mask[u] = True
            assert np.all(dense[~mask] == 0)


@pytest.mark.parametrize("nsrc", [1, 2, 4, 8])
def test_sparse_aggregator_bit_exact(mx, oracle, nsrc, union_path):
    # test_kvstore.py:178-227: push several row_sparse values of one key, pull everything
    rng = np.random.default_rng(nsrc)
    shape = (1000, 48)
    kv = mx.kv.create('device')
    kv.init(9, mx.nd.sparse.zeros('row_sparse', shape, mx.gpu(0)))
    for rep in range(2):
        arrs, idxs, vals = [], [], []
        for _ in range(nsrc):
            a, i, v = make_rsp(mx, rng, shape, 100, mx.gpu(0))
            arrs.append(a)
            idxs.append(i)
            vals.append(v)
        kv.push(9, arrs)
        want_idx, want_val = oracle.rsp_reduce(idxs, vals)
        out = mx.nd.sparse.zeros('row_sparse', shape, mx.gpu(0))
        all_rows = mx.nd.array(np.arange(shape[0], dtype=np.int64), mx.gpu(0), np.int64)
        kv.row_sparse_pull(9, out=out, row_ids=all_rows)
        ri, rv = oracle.sparse_retain(want_idx, want_val, np.arange(shape[0]))
        assert np.array_equal(out.indices.asnumpy(), ri)
        assert eq(out.data.asnumpy(), rv.reshape(out.data.shape))
        # a subset with duplicates, some rows absent from the merged value
        req = rng.integers(0, shape[0], 300)
        out2 = mx.nd.sparse.zeros('row_sparse', shape, mx.gpu(0))
        kv.row_sparse_pull(9, out=out2, row_ids=mx.nd.array(req, mx.gpu(0), np.int64))
        ri, rv = oracle.sparse_retain(want_idx, want_val, oracle.unique(req))
        assert np.array_equal(out2.indices.asnumpy(), ri)
        assert eq(out2.data.asnumpy(), rv.reshape(out2.data.shape))


def test_empty_and_ragged_inputs(mx, oracle, union_path):
    rng = np.random.default_rng(3)
    shape = (50, 5)            # row length not a multiple of 4: scalar row path
    kv = mx.kv.create('device')
    kv.init(0, mx.nd.sparse.zeros('row_sparse', shape, mx.gpu(0)))
    empty = mx.nd.sparse.zeros('row_sparse', shape, mx.gpu(0))
    a, i, v = make_rsp(mx, rng, shape, 7, mx.gpu(0))
    kv.push(0, [empty, a, empty])              # all-zero sources are skipped
    out = mx.nd.sparse.zeros('row_sparse', shape, mx.gpu(0))
    kv.row_sparse_pull(0, out=out, row_ids=mx.nd.array(np.arange(50), mx.gpu(0), np.int64))
    assert eq(out.asnumpy()[i], 0.0 + v)
    kv.push(0, [empty, empty])                 # merged value is empty -> stored becomes all zero
    out = mx.nd.sparse.zeros('row_sparse', shape, mx.gpu(0))
    kv.row_sparse_pull(0, out=out, row_ids=mx.nd.array(np.arange(10), mx.gpu(0), np.int64))
    assert out.indices.shape[0] == 0 and np.all(out.asnumpy() == 0)
    # single row id
    kv.push(0, [a])
    out = mx.nd.sparse.zeros('row_sparse', shape, mx.gpu(0))
    kv.row_sparse_pull(0, out=out, row_ids=mx.nd.array(np.array([int(i[0])]), mx.gpu(0), np.int64))
    assert np.array_equal(out.indices.asnumpy(), [i[0]]) and eq(out.data.asnumpy()[0], v[0])


def test_large_table(mx, oracle, union_path):
    # test_kvstore_gpu.py:126-134 uses a 793470-row table; sizes here keep the oracle in seconds
    rng = np.random.default_rng(4)
    shape = (793470, 16)
    kv = mx.kv.create('device')
    kv.init('big', mx.nd.sparse.zeros('row_sparse', shape, mx.gpu(0)))
    srcs = [make_rsp(mx, rng, shape, 5000, mx.gpu(0)) for _ in range(4)]
    kv.push('big', [s[0] for s in srcs])
    want_idx, want_val = oracle.rsp_reduce([s[1] for s in srcs], [s[2] for s in srcs])
    req = np.concatenate([want_idx[::3], rng.integers(0, shape[0], 2000)])
    out = mx.nd.sparse.zeros('row_sparse', shape, mx.gpu(0))
    kv.row_sparse_pull('big', out=out, row_ids=mx.nd.array(req, mx.gpu(0), np.int64))
    ri, rv = oracle.sparse_retain(want_idx, want_val, oracle.unique(req))
    assert np.array_equal(out.indices.asnumpy(), ri)
    assert eq(out.data.asnumpy(), rv)


@pytest.mark.parametrize("optname,clip", [('sgd', None), ('sgd_mom', 0.3), ('adam', None), ('adam', 0.3)])
def test_fused_lazy_sparse_update(mx, oracle, optname, clip, union_path):
    """Embedding-style key: weight holds every row, gradients are row_sparse, optimizer on the store
    (trainer.py:177-191 forces update_on_kvstore for sparse parameters)."""
    rng = np.random.default_rng(5)
    shape = (400, 32)
    w = rng.uniform(-1, 1, shape).astype(np.float32)
    kv = mx.kv.create('device')
    kv.init(0, mx.nd.array(w, mx.gpu(0)).tostype('row_sparse'))
    if optname.startswith('sgd'):
        mom = 0.9 if optname == 'sgd_mom' else 0.0
        kv.set_optimizer(mx.optimizer.SGD(learning_rate=0.1, momentum=mom, wd=1e-3, rescale_grad=0.5,
                                          clip_gradient=clip))
    else:
        kv.set_optimizer(mx.optimizer.Adam(learning_rate=1e-3, wd=0.01, rescale_grad=0.5,
                                           clip_gradient=clip))
    sp = K.scalar_param
    m = np.zeros(shape, np.float32)
    v = np.zeros(shape, np.float32)
    for t in range(1, 4):
        srcs = [make_rsp(mx, rng, shape, 60, mx.gpu(0)) for _ in range(3)]
        kv.push(0, [s[0] for s in srcs])
        gi, gv = oracle.rsp_reduce([s[1] for s in srcs], [s[2] for s in srcs])
        if optname == 'sgd':
            oracle.sgd_rsp_update(w, gi, gv, sp(0.1), sp(1e-3), sp(0.5), sp(clip) if clip else None)
        elif optname == 'sgd_mom':
            oracle.sgd_mom_rsp_update(w, m, gi, gv, sp(0.1), sp(0.9), sp(1e-3), sp(0.5),
                                      sp(clip) if clip else None)
        else:
            oracle.adam_rsp_update(w, m, v, gi, gv, sp(K.adam_lr(1e-3, 0.9, 0.999, t)), sp(0.9),
                                   sp(0.999), sp(1e-8), sp(0.01), sp(0.5), sp(clip) if clip else None)
        out = mx.nd.sparse.zeros('row_sparse', shape, mx.gpu(0))
        kv.row_sparse_pull(0, out=out, row_ids=mx.nd.array(np.arange(shape[0]), mx.gpu(0), np.int64))
        assert eq(out.asnumpy(), w), t


@pytest.mark.parametrize("tag,clip", [("noclip", None), ("clip", 0.3)])
def test_lazy_update_ops_golden(mx, golden, tag, clip):
    """sgd_update / sgd_mom_update / adam_update with a row_sparse gradient vs the reference's own
    <req,cpu> kernels (tests/golden/rowsparse_updates.npz)."""
    g = golden("rowsparse_updates")
    dev = mx.gpu(0)
    w, m, v = g["rsp_%s_in" % tag]
    gi, gv = g["rsp_%s_gidx" % tag], g["rsp_%s_gval" % tag]
    grad = mx.nd.sparse.row_sparse_array((gv, gi), shape=w.shape, ctx=dev)
    kw = dict(clip_gradient=clip) if clip else {}
    wn = mx.nd.array(w, dev)
    mx.nd.sgd_update(wn, grad, out=wn, lr=0.1, wd=1e-3, rescale_grad=0.5, lazy_update=True, **kw)
    assert eq(wn.asnumpy(), g["rsp_sgd_%s_out" % tag])
    wn, mn = mx.nd.array(w, dev), mx.nd.array(m, dev)
    mx.nd.sgd_mom_update(wn, grad, mn, out=wn, lr=0.1, momentum=0.9, wd=1e-3, rescale_grad=0.5,
                         lazy_update=True, **kw)
    assert eq(np.stack([wn.asnumpy(), mn.asnumpy()]), g["rsp_sgdmom_%s_out" % tag])
    wn, mn, vn = mx.nd.array(w, dev), mx.nd.array(m, dev), mx.nd.array(v, dev)
    mx.nd.adam_update(wn, grad, mn, vn, out=wn, lr=1e-3, wd=0.01, lazy_update=True, **kw)
    got = np.stack([wn.asnumpy(), mn.asnumpy(), vn.asnumpy()])
    # the fixture used the C float defaults (0.9f, 0.999f, 1e-8f, 0.001f, 0.01f) == what
    # dmlc::stof yields for those literals
    assert eq(got, g["rsp_adam_%s_out" % tag])


def test_host_resident_sparse_values(mx, oracle):
    rng = np.random.default_rng(6)
    shape = (200, 12)
    kv = mx.kv.create('local')
    kv.init(1, mx.nd.sparse.zeros('row_sparse', shape, mx.cpu()))
    srcs = [make_rsp(mx, rng, shape, 30, mx.cpu(), integer=True) for _ in range(3)]
    kv.push(1, [s[0] for s in srcs])
    out = mx.nd.sparse.zeros('row_sparse', shape, mx.cpu())
    kv.row_sparse_pull(1, out=out, row_ids=mx.nd.array(np.arange(shape[0]), mx.cpu(), np.int64))
    dense = np.zeros(shape, np.float32)
    for _, i, v in srcs:
        dense[i] += v
    assert np.array_equal(out.asnumpy(), dense)     # small integers: exact


# ---------------------------------------------------------------------------------- SURVEY 8f-f3
@pytest.mark.parametrize("tag,clip", [("noclip", None), ("clip", 0.3)])
def test_std_update_ops_golden(mx, golden, tag, clip):
    """lazy_update=False: every row moves. Against the reference's own <req,cpu> std kernels
    (optimizer_op.cc:108-139, 195-229; tests/golden/rowsparse_std_updates.npz), bit for bit."""
    g = golden("rowsparse_std_updates")
    dev = mx.gpu(0)
    w, m, v = g["std_%s_in" % tag]
    gi, gv = g["std_%s_gidx" % tag], g["std_%s_gval" % tag]
    grad = mx.nd.sparse.row_sparse_array((gv, gi), shape=w.shape, ctx=dev)
    kw = dict(clip_gradient=clip) if clip else {}
    wn = mx.nd.array(w, dev)
    mx.nd.sgd_update(wn, grad, out=wn, lr=0.1, wd=1e-3, rescale_grad=0.5, lazy_update=False, **kw)
    assert eq(wn.asnumpy(), g["std_sgd_%s_out" % tag])
    wn, mn = mx.nd.array(w, dev), mx.nd.array(m, dev)
    mx.nd.sgd_mom_update(wn, grad, mn, out=wn, lr=0.1, momentum=0.9, wd=1e-3, rescale_grad=0.5,
                         lazy_update=False, **kw)
    assert eq(np.stack([wn.asnumpy(), mn.asnumpy()]), g["std_sgdmom_%s_out" % tag])
    wn, mn, vn = mx.nd.array(w, dev), mx.nd.array(m, dev), mx.nd.array(v, dev)
    mx.nd.adam_update(wn, grad, mn, vn, out=wn, lr=1e-3, wd=0.01, lazy_update=False, **kw)
    assert eq(np.stack([wn.asnumpy(), mn.asnumpy(), vn.asnumpy()]), g["std_adam_%s_out" % tag])
    # an all-zero gradient still decays / moves every row
    wn, mn = mx.nd.array(w, dev), mx.nd.array(m, dev)
    empty = mx.nd.sparse.zeros('row_sparse', w.shape, dev)
    mx.nd.sgd_mom_update(wn, empty, mn, out=wn, lr=0.1, momentum=0.9, wd=1e-3, rescale_grad=0.5,
                         lazy_update=False, **kw)
    assert eq(np.stack([wn.asnumpy(), mn.asnumpy()]), g["std_sgdmom_%s_empty_out" % tag])


@pytest.mark.parametrize("optname", ['sgd', 'sgd_mom', 'adam'])
def test_store_std_sparse_update(mx, oracle, optname):
    """optimizer on the store with lazy_update=False and row_sparse pushes"""
    rng = np.random.default_rng(15)
    shape = (300, 20)
    w = rng.uniform(-1, 1, shape).astype(np.float32)
    kv = mx.kv.create('device')
    kv.init(0, mx.nd.array(w, mx.gpu(0)).tostype('row_sparse'))
    if optname.startswith('sgd'):
        mom = 0.9 if optname == 'sgd_mom' else 0.0
        kv.set_optimizer(mx.optimizer.SGD(learning_rate=0.1, momentum=mom, wd=1e-3, rescale_grad=0.5,
                                          lazy_update=False))
    else:
        kv.set_optimizer(mx.optimizer.Adam(learning_rate=1e-3, wd=0.01, rescale_grad=0.5,
                                           lazy_update=False))
    sp = K.scalar_param
    m, v = np.zeros(shape, np.float32), np.zeros(shape, np.float32)
    for t in range(1, 4):
        srcs = [make_rsp(mx, rng, shape, 40, mx.gpu(0)) for _ in range(3)]
        kv.push(0, [s[0] for s in srcs])
        gi, gv = oracle.rsp_reduce([s[1] for s in srcs], [s[2] for s in srcs])
        if optname == 'sgd':
            oracle.sgd_std_rsp_update(w, gi, gv, sp(0.1), sp(1e-3), sp(0.5))
        elif optname == 'sgd_mom':
            oracle.sgd_mom_std_rsp_update(w, m, gi, gv, sp(0.1), sp(0.9), sp(1e-3), sp(0.5))
        else:
            oracle.adam_std_rsp_update(w, m, v, gi, gv, sp(K.adam_lr(1e-3, 0.9, 0.999, t)), sp(0.9),
                                       sp(0.999), sp(1e-8), sp(0.01), sp(0.5))
        out = mx.nd.sparse.zeros('row_sparse', shape, mx.gpu(0))
        kv.row_sparse_pull(0, out=out, row_ids=mx.nd.array(np.arange(shape[0]), mx.gpu(0), np.int64))
        assert eq(out.asnumpy(), w), t


def test_pull_with_storage_casts(mx):
    """pull(ignore_sparse=False) (kvstore.py:233-319; comm Broadcast -> CopyFromTo casts,
    ndarray.cc:1147-1196, cast_storage-inl.h:74-140)"""
    rng = np.random.default_rng(21)
    shape = (50, 8)
    dense = rng.uniform(-1, 1, shape).astype(np.float32)
    dense[[0, 7, 8, 33, 49]] = 0          # dense -> row_sparse drops all-zero rows
    dense[5, 3] = -0.0                    # ... a negative zero is still zero
    dense[5, :3] = 0
    dense[5, 4:] = 0
    kv = mx.kv.create('device')
    kv.init('d', mx.nd.array(dense, mx.gpu(0)))
    out = mx.nd.sparse.zeros('row_sparse', shape, mx.gpu(0))
    kv.pull('d', out=out, ignore_sparse=False)
    keep = np.array([r for r in range(shape[0]) if np.any(dense[r] != 0)], np.int64)
    assert np.array_equal(out.indices.asnumpy(), keep)
    assert eq(out.data.asnumpy(), dense[keep])
    # ignore_sparse=True (default): a row_sparse target of a pull is skipped
    untouched = mx.nd.sparse.zeros('row_sparse', shape, mx.gpu(0))
    kv.pull('d', out=untouched)
    assert untouched.indices.shape[0] == 0
    # row_sparse key: full copy into a row_sparse target, scatter into a dense one
    idx = np.array([2, 3, 11, 48], np.int64)
    val = rng.uniform(-1, 1, (4, 8)).astype(np.float32)
    kv.init('r', mx.nd.sparse.row_sparse_array((val, idx), shape=shape, ctx=mx.gpu(0)))
    out_r = mx.nd.sparse.zeros('row_sparse', shape, mx.gpu(0))
    kv.pull('r', out=out_r, ignore_sparse=False)
    assert np.array_equal(out_r.indices.asnumpy(), idx) and eq(out_r.data.asnumpy(), val)
    out_d = mx.nd.array(np.full(shape, 9, np.float32), mx.gpu(0))
    kv.pull('r', out=out_d, ignore_sparse=False)
    want = np.zeros(shape, np.float32)
    want[idx] = val
    assert eq(out_d.asnumpy(), want)
    # copyto across storage types and devices (host target)
    host = mx.nd.zeros(shape, mx.cpu())
    out_r.copyto(host)
    assert eq(host.asnumpy(), want)
