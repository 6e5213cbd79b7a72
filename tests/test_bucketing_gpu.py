"""Deferred bucket execution (B200KVStoreSetBucketBytes): the reference's per-parameter call pattern
(gluon/trainer.py:371-396: pushpull(i, grads, out=weights, priority=-i); tools/bandwidth/measure.py:
push(i) then pull(i)) is queued and fused into few launches, with unchanged results."""
import ctypes

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


def set_bucket(mx, kv, nbytes):
    mx.base.check_call(mx.base._LIB.B200KVStoreSetBucketBytes(kv.handle, ctypes.c_size_t(nbytes)))


@pytest.mark.parametrize("pattern", ["pushpull", "push_then_pull"])
def test_per_key_calls_are_fused(mx, pattern):
    rng = np.random.default_rng(30)
    shapes = [(64, 3, 7, 7), (64,), (3,), (256, 64, 1, 1), (1000, 2048), (1000,), (512, 512, 3, 3)]
    kv = mx.kv.create('device')
    model = K.LocalKVStoreModel('device')
    for k, s in enumerate(shapes):
        w = rng.uniform(-1, 1, s).astype(np.float32)
        kv.init(k, mx.nd.array(w, mx.gpu(0)))
        model.init(k, w)
    kv.set_optimizer(mx.optimizer.SGD(learning_rate=0.1, momentum=0.9, wd=1e-4, rescale_grad=1 / 32))
    model.set_optimizer('sgd', lr=0.1, momentum=0.9, wd=1e-4, rescale_grad=1 / 32)
    set_bucket(mx, kv, 1 << 30)       # large bucket: everything queued until something waits
    outs = [mx.nd.empty(s, mx.gpu(0)) for s in shapes]
    for step in range(3):
        grads = [[rng.uniform(-1, 1, s).astype(np.float32) for _ in range(2)] for s in shapes]
        vals = [[mx.nd.array(g, mx.gpu(0)) for g in gs] for gs in grads]
        mx.nd.waitall()
        before = mx.base.kernel_launch_count()
        for k in range(len(shapes)):
            if pattern == "pushpull":
                kv.pushpull(k, vals[k], out=outs[k], priority=-k)
            else:
                kv.push(k, vals[k], priority=-k)
                kv.pull(k, out=outs[k], priority=-k)
        assert mx.base.kernel_launch_count() == before      # nothing launched yet: queued
        outs[0].wait_to_read()                                # any wait flushes the queue
        assert mx.base.kernel_launch_count() - before == 1  # ONE fused launch for all keys
        for k in range(len(shapes)):
            model.push(k, grads[k])
            assert eq(outs[k].asnumpy(), model.pull(k)), (step, k)


def test_bucket_threshold_and_repeated_key(mx, oracle):
    rng = np.random.default_rng(31)
    n = 100000                                     # 400 KB per value
    kv = mx.kv.create('local')
    for k in range(6):
        kv.init(k, mx.nd.zeros((n,), mx.gpu(0)))
    set_bucket(mx, kv, 1 << 20)                    # 1 MB: a flush every ~2 keys (value + out)
    srcs = [rng.uniform(-1, 1, n).astype(np.float32) for _ in range(6)]
    outs = [mx.nd.empty((n,), mx.gpu(0)) for _ in range(6)]
    vals = [mx.nd.array(s, mx.gpu(0)) for s in srcs]    # (creating arrays observes memory: flushes)
    mx.nd.waitall()
    before = mx.base.kernel_launch_count()
    for k in range(6):
        kv.pushpull(k, vals[k], out=outs[k])
    mx.nd.waitall()
    launches = mx.base.kernel_launch_count() - before
    assert 2 <= launches <= 4
    for k in range(6):
        assert eq(outs[k].asnumpy(), srcs[k])
    # two pushes of the SAME key must stay two updates (no merging into one reduce)
    kv2 = mx.kv.create('local')
    kv2.init(0, mx.nd.zeros((n,), mx.gpu(0)))
    kv2.set_optimizer(mx.optimizer.Test(rescale_grad=1.0))
    set_bucket(mx, kv2, 1 << 30)
    a, b = srcs[0], srcs[1]
    out = mx.nd.empty((n,), mx.gpu(0))
    kv2.push(0, vals[0])
    kv2.push(0, vals[1])
    kv2.pull(0, out=out)
    assert eq(out.asnumpy(), (np.zeros(n, np.float32) + a) + b)
    # errors stay synchronous even when queueing
    with pytest.raises(mx.MXNetError, match="has not been inited"):
        kv2.push(7, mx.nd.zeros((n,), mx.gpu(0)))
    with pytest.raises(mx.MXNetError, match="shape mismatch"):
        kv2.push(0, mx.nd.zeros((3,), mx.gpu(0)))


def test_foreign_host_memory_is_staged(mx, oracle):
    """Host tensors the library did not allocate (torch CPU memory through DLPack) cannot be read by
    the kernel directly: they take the copy-lane staging path (H2D / D2H DMA in buckets)."""
    import torch
    rng = np.random.default_rng(32)
    shapes = [(300, 1000), (17,), (2048, 1024)]
    kv = mx.kv.create('device')
    keys = list(range(len(shapes)))
    kv.init(keys, [mx.nd.zeros(s, mx.gpu(0)) for s in shapes])
    srcs = [[rng.uniform(-1, 1, s).astype(np.float32) for _ in range(2)] for s in shapes]
    tin = [[torch.from_numpy(a).pin_memory() for a in ss] for ss in srcs]
    tout = [torch.zeros(s).pin_memory() for s in shapes]
    kv.pushpull(keys, [[mx.nd.from_torch(t) for t in ts] for ts in tin],
                out=[mx.nd.from_torch(t) for t in tout])
    mx.nd.waitall()
    for k in keys:
        assert eq(tout[k].numpy().ravel(), oracle.reduce(srcs[k], 'device'))
