"""Parity at the FULL sizes BASELINE.json's configs name, on ONE GPU (so the driver's single-GPU
box runs them), through the C ABI, against the CPU oracle:

  cfg1  kvstore('local'), one fp32 key of shape (1024, 1024), CPU-context values
  cfg2  ResNet-50 gradient set (157 tensors / 25 549 486 elements), SGD-momentum fused on the store
  cfg4  BERT-base gradient set (199 tensors / 109 482 240 elements, 4 values per key), Adam fused
  cfg5  row_sparse table (1 000 000, 512), 8 values x 10 000 rows, pull with 5 % duplicate ids

Shapes of the checks follow tests/nightly/test_kvstore.py:79-98,214-258 (random data, optimizer on
the store, several steps) and tests/python/gpu/test_kvstore_gpu.py:46-134 (row_sparse push / pull) of
the reference; comparisons are BIT-EXACT (stronger than the reference's 1e-6 relative bound).
"""
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


def flat_set(seed, sizes):
    """one rng stream per (role, source): element values of every tensor of the set"""
    rng = np.random.default_rng(seed)
    flat = rng.uniform(-1, 1, int(sum(sizes))).astype(np.float32)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    return [flat[offs[i]:offs[i + 1]] for i in range(len(sizes))]


def test_cfg1_local_single_key_cpu_values(mx):
    # BASELINE configs[0]; tests/python/unittest/test_kvstore.py:38-66 with the config's shape
    shape = (1024, 1024)
    rng = np.random.default_rng(1)
    kv = mx.kv.create('local')
    model = K.LocalKVStoreModel('local')
    w = rng.uniform(-1, 1, shape).astype(np.float32)
    kv.init(3, mx.nd.array(w, mx.cpu()))
    model.init(3, w)
    for n in (1, 4):
        vals = [rng.uniform(-1, 1, shape).astype(np.float32) for _ in range(n)]
        kv.push(3, [mx.nd.array(v, mx.cpu()) for v in vals])
        model.push(3, vals)
        out = mx.nd.empty(shape, mx.cpu())
        kv.pull(3, out=out)
        assert eq(out.asnumpy(), model.pull(3))


def run_dense_config(mx, shapes, nsrc, opt_name, opt_kw, steps):
    sizes = [int(np.prod(s)) for s in shapes]
    keys = list(range(len(shapes)))
    kv = mx.kv.create('device')
    model = K.LocalKVStoreModel('device')
    w0 = flat_set(0xB200 + 777, sizes)
    kv.init(keys, [mx.nd.array(w0[k].reshape(shapes[k]), mx.gpu(0)) for k in keys])
    for k in keys:
        model.init(k, w0[k].reshape(shapes[k]))
    kv.set_optimizer(getattr(mx.optimizer, opt_name)(**opt_kw))
    mkw = dict(opt_kw)
    mkw['lr'] = mkw.pop('learning_rate')
    model.set_optimizer(opt_name.lower(), **mkw)
    outs = [mx.nd.empty(s, mx.gpu(0)) for s in shapes]
    before = mx.base.kernel_launch_count()
    for t in range(steps):
        g = [flat_set(0xB200 + 1000 * j + 17 * t, sizes) for j in range(nsrc)]
        vals = [[mx.nd.array(g[j][k].reshape(shapes[k]), mx.gpu(0)) for j in range(nsrc)] for k in keys]
        kv.pushpull(keys, vals if nsrc > 1 else [v[0] for v in vals], out=outs)
        for k in keys:
            model.push(k, [g[j][k] for j in range(nsrc)])
        del vals
    mx.nd.waitall()
    # the whole set is ONE fused launch per step (plus the staging of mx.nd.array uploads, which
    # the engine does not count as kernels)
    assert mx.base.kernel_launch_count() - before == steps
    bad = [k for k in keys if not eq(outs[k].asnumpy(), model.pull(k))]
    assert not bad, "keys differing from the oracle: %s" % bad[:10]
    # a later plain pull returns the same bits (idempotent)
    outs2 = [mx.nd.empty(s, mx.gpu(0)) for s in shapes]
    kv.pull(keys, out=outs2)
    for k in (0, len(keys) // 2, len(keys) - 1):
        assert eq(outs2[k].asnumpy(), model.pull(k))


def test_cfg2_resnet50_set_sgd_momentum_full(mx):
    from bench import resnet50_shapes, SGD_KW
    shapes = resnet50_shapes()
    assert len(shapes) == 157 and sum(int(np.prod(s)) for s in shapes) == 25549486
    run_dense_config(mx, shapes, 1, 'SGD', dict(rescale_grad=1.0 / 256, **SGD_KW), steps=3)


def test_cfg2_resnet50_set_two_values_clip(mx):
    # two values per key (the reduce is in the loop) and gradient clipping on ("clip off
    # and 1.0, two runs")
    from bench import resnet50_shapes, SGD_KW
    run_dense_config(mx, resnet50_shapes(), 2, 'SGD',
                     dict(rescale_grad=1.0 / 2, clip_gradient=1.0, **SGD_KW), steps=2)


def test_cfg4_bert_base_set_adam_full(mx):
    from bench import bert_base_shapes, ADAM_KW
    shapes = bert_base_shapes()
    assert len(shapes) == 199 and sum(int(np.prod(s)) for s in shapes) == 109482240
    run_dense_config(mx, shapes, 4, 'Adam', dict(rescale_grad=1.0 / 4, **ADAM_KW), steps=2)


def test_cfg5_row_sparse_1m_x_512(mx, oracle):
    rows, row_len, hot, nval = 1000000, 512, 10000, 8
    shape = (rows, row_len)
    rng = np.random.default_rng(0xB200)
    w = rng.uniform(-1, 1, shape).astype(np.float32)
    kv = mx.kv.create('device')
    kv.init('emb', mx.nd.sparse.row_sparse_array((w, np.arange(rows, dtype=np.int64)), shape=shape,
                                                 ctx=mx.gpu(0)))
    lr, rescale = 0.1, 1.0 / nval
    kv.set_optimizer(mx.optimizer.SGD(learning_rate=lr, momentum=0.0, wd=0.0, rescale_grad=rescale))
    for step in range(2):
        idx = [np.sort(rng.choice(rows, hot, replace=False)).astype(np.int64) for _ in range(nval)]
        val = [rng.uniform(-1, 1, (hot, row_len)).astype(np.float32) for _ in range(nval)]
        kv.push('emb', [mx.nd.sparse.row_sparse_array((val[i], idx[i]), shape=shape, ctx=mx.gpu(0))
                        for i in range(nval)])
        # oracle: union + in-order accumulation (ndarray_function.cc:59-175), then the lazy
        # sgd_update on the rows of the merged gradient (optimizer_op-inl.h:426-450)
        gi, gv = oracle.rsp_reduce(idx, val)
        assert 70000 < len(gi) <= nval * hot
        oracle.sgd_rsp_update(w, gi, gv, K.scalar_param(lr), K.scalar_param(0.0), K.scalar_param(rescale), None)
    # every "GPU" pulls its own id list: unsorted, 5 % duplicates (exercises Unique)
    pull_ids, outs = [], []
    for i in range(nval):
        ids = np.concatenate([idx[i], rng.choice(idx[i], hot // 20)])
        rng.shuffle(ids)
        pull_ids.append(ids)
        outs.append(mx.nd.sparse.zeros('row_sparse', shape, mx.gpu(0)))
    kv.row_sparse_pull('emb', out=outs, row_ids=[mx.nd.array(p, mx.gpu(0), np.int64) for p in pull_ids])
    for i in range(nval):
        u = oracle.unique(pull_ids[i])
        assert np.array_equal(outs[i].indices.asnumpy(), u)
        assert eq(outs[i].data.asnumpy(), w[u])
    # rows that no gradient touched are unchanged, bit for bit (pull a sample of cold rows)
    cold = np.setdiff1d(np.arange(0, rows, 997), np.concatenate(idx))
    out = mx.nd.sparse.zeros('row_sparse', shape, mx.gpu(0))
    kv.row_sparse_pull('emb', out=out, row_ids=mx.nd.array(cold, mx.gpu(0), np.int64))
    assert np.array_equal(out.indices.asnumpy(), cold) and eq(out.data.asnumpy(), w[cold])
