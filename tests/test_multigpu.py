"""Multi-GPU (single process, peer access -- the reference's own process model) parity tests.
Modeled on tests/python/gpu/test_device.py and tests/nightly/test_kvstore.py. Needs >= 2 GPUs."""
import numpy as np
import pytest

import kvoracle as K

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


@pytest.fixture(scope="module")
def mx():
    import anand_mxnet_b200 as mx
    return mx


def ngpu():
    import torch
    return torch.cuda.device_count()


def eq(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def rnd(rng, shape):
    return rng.uniform(-1, 1, shape).astype(np.float32)


SHAPES = [(4, 4), (100, 100), (3,), (1027,), (70001, 3), (2000, 2000)]


@pytest.mark.parametrize("kvtype", ['local', 'device'])
def test_ones_over_devices(mx, kvtype):
    # test_device.py:38-71
    n = ngpu()
    kv = mx.kv.create(kvtype)
    shapes = [(10,), (100, 50), (2, 3, 4, 5, 6, 7, 8)]
    for k, s in enumerate(shapes):
        kv.init(k, mx.nd.zeros(s, mx.gpu(0)))
    for k, s in enumerate(shapes):
        vals = [mx.nd.ones(s, mx.gpu(d)) for d in range(n)]
        outs = [mx.nd.empty(s, mx.gpu(d)) for d in range(n)]
        kv.push(k, vals)
        kv.pull(k, out=outs)
        for o in outs:
            assert np.all(o.asnumpy() == n)


@pytest.mark.parametrize("kvtype,order", [('local', 'local'), ('device', 'device')])
def test_reduce_and_broadcast_bit_exact(mx, oracle, kvtype, order):
    n = ngpu()
    rng = np.random.default_rng(20)
    kv = mx.kv.create(kvtype)
    keys = list(range(len(SHAPES)))
    kv.init(keys, [mx.nd.zeros(s, mx.gpu(0)) for s in SHAPES])
    for rep in range(2):
        srcs = [[rnd(rng, s) for _ in range(n)] for s in SHAPES]
        vals = [[mx.nd.array(a, mx.gpu(d)) for d, a in enumerate(srcs[k])] for k in keys]
        outs = [[mx.nd.empty(s, mx.gpu(d)) for d in range(n)] for s in SHAPES]
        kv.pushpull(keys, vals, out=outs)
        for k in keys:
            want = oracle.reduce(srcs[k], order)
            for o in outs[k]:
                assert eq(o.asnumpy().ravel(), want), (rep, k)


@pytest.mark.parametrize("optname", ['sgd', 'adam'])
def test_fused_optimizer_striped(mx, optname):
    """Stripes of the key space are owned by different GPUs; the result must not depend on that."""
    n = ngpu()
    rng = np.random.default_rng(21)
    kv = mx.kv.create('device')
    model = K.LocalKVStoreModel('device')
    keys = list(range(len(SHAPES)))
    for k, s in enumerate(SHAPES):
        w = rnd(rng, s)
        kv.init(k, mx.nd.array(w, mx.gpu(k % n)))      # initial values scattered over the GPUs
        model.init(k, w)
    if optname == 'sgd':
        kv.set_optimizer(mx.optimizer.SGD(learning_rate=0.1, momentum=0.9, wd=1e-4,
                                          rescale_grad=1.0 / (64 * n), clip_gradient=0.01))
        model.set_optimizer('sgd', lr=0.1, momentum=0.9, wd=1e-4, rescale_grad=1.0 / (64 * n),
                            clip_gradient=0.01)
    else:
        kv.set_optimizer(mx.optimizer.Adam(learning_rate=1e-3, wd=0.01))
        model.set_optimizer('adam', lr=1e-3, wd=0.01)
    outs = [[mx.nd.empty(s, mx.gpu(d)) for d in range(n)] for s in SHAPES]
    for step in range(3):
        grads = [[rnd(rng, s) for _ in range(n)] for s in SHAPES]
        vals = [[mx.nd.array(g, mx.gpu(d)) for d, g in enumerate(grads[k])] for k in keys]
        kv.pushpull(keys, vals, out=outs)
        for k in keys:
            model.push(k, grads[k])
            want = model.pull(k)
            for o in outs[k]:
                assert eq(o.asnumpy(), want), (step, k)
    # a later plain pull (all-gather of the owners' stripes) returns the same weights
    fresh = [mx.nd.empty(s, mx.gpu(n - 1)) for s in SHAPES]
    kv.pull(keys, out=fresh)
    for k in keys:
        assert eq(fresh[k].asnumpy(), model.pull(k))
    # optimizer-state checkpoint gathers the stripes back
    import tempfile, os, pickle
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "states")
        kv.save_optimizer_states(f)
        states = pickle.load(open(f, "rb"))
        st = states[5]
        got = (st[0] if isinstance(st, tuple) else st).asnumpy()
        want = model.state[5][0] if optname == 'adam' else model.state[5]
        assert eq(got, want)


def test_updater_callback_multi_device(mx):
    # nightly test_kvstore.py: 'test' optimizer through the Python updater over several devices
    n = ngpu()

    def updater(key, recv, local):
        local += recv
    kv = mx.kv.create('device')
    kv._set_updater(updater)
    kv.init(3, mx.nd.ones((50, 50), mx.gpu(0)))
    for _ in range(3):
        kv.push(3, [mx.nd.ones((50, 50), mx.gpu(d)) for d in range(n)])
    outs = [mx.nd.empty((50, 50), mx.gpu(d)) for d in range(n)]
    kv.pull(3, out=outs)
    for o in outs:
        assert np.all(o.asnumpy() == 1 + 3 * n)


def test_row_sparse_multi_device(mx, oracle):
    # test_kvstore_gpu.py:46-110: rsp values pushed from several GPUs, pulled to several GPUs
    n = ngpu()
    rng = np.random.default_rng(22)
    shape = (5000, 64)
    kv = mx.kv.create('device')
    kv.init('e', mx.nd.sparse.zeros('row_sparse', shape, mx.gpu(0)))
    idxs, vals, arrs = [], [], []
    for d in range(n):
        i = np.sort(rng.choice(shape[0], 300, replace=False)).astype(np.int64)
        v = rnd(rng, (300, shape[1]))
        idxs.append(i)
        vals.append(v)
        arrs.append(mx.nd.sparse.row_sparse_array((v, i), shape=shape, ctx=mx.gpu(d)))
    kv.push('e', arrs)
    wi, wv = oracle.rsp_reduce(idxs, vals)
    outs = [mx.nd.sparse.zeros('row_sparse', shape, mx.gpu(d)) for d in range(n)]
    ids = [rng.integers(0, shape[0], 500) for _ in range(n)]
    kv.row_sparse_pull('e', out=outs, row_ids=[mx.nd.array(i, mx.gpu(d), np.int64) for d, i in enumerate(ids)])
    for o, i in zip(outs, ids):
        ri, rv = oracle.sparse_retain(wi, wv, oracle.unique(i))
        assert np.array_equal(o.indices.asnumpy(), ri)
        assert eq(o.data.asnumpy(), rv)


@pytest.mark.parametrize("optname", ['sgd_mom', 'adam'])
def test_row_sparse_table_sharded_by_row_range(optname):
    """gradients pushed from two GPUs -> every GPU merges and updates its own row range
    of the table (weight + optimizer state sharded); pulls gather rows from the owning shards.
    Bit-exact against the oracle, before and after folding the shards back."""
    import kvoracle as K
    import anand_mxnet_b200 as mx
    oracle = K.get_oracle()
    rng = np.random.default_rng(31)
    shape = (1001, 24)                      # odd height: unequal shards
    w = rng.uniform(-1, 1, shape).astype(np.float32)
    ctxs = [mx.gpu(0), mx.gpu(1)]
    kv = mx.kv.create('device')
    kv.init('emb', mx.nd.array(w, ctxs[0]).tostype('row_sparse'))
    if optname == 'sgd_mom':
        kv.set_optimizer(mx.optimizer.SGD(learning_rate=0.1, momentum=0.9, wd=1e-3, rescale_grad=0.5))
    else:
        kv.set_optimizer(mx.optimizer.Adam(learning_rate=1e-3, wd=0.01, rescale_grad=0.5))
    sp = K.scalar_param
    m, v = np.zeros(shape, np.float32), np.zeros(shape, np.float32)

    def make(ctx, n):
        idx = np.sort(rng.choice(shape[0], n, replace=False)).astype(np.int64)
        val = rng.uniform(-1, 1, (n, shape[1])).astype(np.float32)
        return mx.nd.sparse.row_sparse_array((val, idx), shape=shape, ctx=ctx), idx, val

    def check(t):
        for c in ctxs:
            ids = rng.integers(0, shape[0], 300)
            out = mx.nd.sparse.zeros('row_sparse', shape, c)
            kv.row_sparse_pull('emb', out=out, row_ids=mx.nd.array(ids, c, np.int64))
            u = np.unique(ids)
            assert np.array_equal(out.indices.asnumpy(), u), t
            assert np.array_equal(out.data.asnumpy().view(np.uint32), w[u].view(np.uint32)), t

    for t in range(1, 5):
        srcs = [make(ctxs[i % 2], 90) for i in range(4)]      # two values per GPU
        kv.push('emb', [s[0] for s in srcs])
        gi, gv = oracle.rsp_reduce([s[1] for s in srcs], [s[2] for s in srcs])
        if optname == 'sgd_mom':
            oracle.sgd_mom_rsp_update(w, m, gi, gv, sp(0.1), sp(0.9), sp(1e-3), sp(0.5), None)
        else:
            oracle.adam_rsp_update(w, m, v, gi, gv, sp(K.adam_lr(1e-3, 0.9, 0.999, t)), sp(0.9), sp(0.999),
                                   sp(1e-8), sp(0.01), sp(0.5), None)
        check(t)
        if t == 2:
            # whole-value access folds the shards back onto one GPU; the next push shards again
            full = mx.nd.sparse.zeros('row_sparse', shape, ctxs[1])
            kv.pull('emb', out=full, ignore_sparse=False)
            assert np.array_equal(full.asnumpy().view(np.uint32), w.view(np.uint32))
    # a push from a single GPU takes the single-owner path on the folded table
    s0 = make(ctxs[0], 50)
    kv.push('emb', s0[0])
    if optname == 'sgd_mom':
        oracle.sgd_mom_rsp_update(w, m, s0[1], s0[2], sp(0.1), sp(0.9), sp(1e-3), sp(0.5), None)
    else:
        oracle.adam_rsp_update(w, m, v, s0[1], s0[2], sp(K.adam_lr(1e-3, 0.9, 0.999, 5)), sp(0.9), sp(0.999),
                               sp(1e-8), sp(0.01), sp(0.5), None)
    check(5)
