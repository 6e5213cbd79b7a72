"""2-bit gradient compression with residual on the device store: reference
semantics from src/kvstore/comm.h:552-596 + gradient_compression-inl.h:40-132, pinned by the
reference's own kernels through tests/golden/twobit.npz; the nightly test
(tests/nightly/test_kvstore.py:100-212) is the model for the multi-step residual checks."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mx():
    import anand_mxnet_b200 as mx
    return mx


def eq(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_twobit_golden_through_store(mx, golden):
    g = golden("twobit")
    grad = g["grad"]
    kv = mx.kv.create('device')
    kv.set_gradient_compression({'type': '2bit', 'threshold': 0.5})
    kv.init(0, mx.nd.zeros(grad.shape, mx.gpu(0)))
    out = mx.nd.empty(grad.shape, mx.gpu(0))
    kv.push(0, mx.nd.array(grad, mx.gpu(0)))
    kv.pull(0, out=out)
    assert eq(out.asnumpy(), g["deq1"])                      # decode of the reference's comp1


@pytest.mark.parametrize("nsrc", [1, 2, 4])
def test_twobit_residual_over_steps(mx, oracle, nsrc):
    rng = np.random.default_rng(50 + nsrc)
    shapes = [(1003,), (64, 33), (16,), (5,)]
    t = 0.25
    kv = mx.kv.create('device')
    kv.set_gradient_compression({'type': '2bit', 'threshold': t})
    keys = list(range(len(shapes)))
    kv.init(keys, [mx.nd.zeros(s, mx.gpu(0)) for s in shapes])
    res = [[np.zeros(int(np.prod(s)), np.float32) for _ in range(nsrc)] for s in shapes]
    outs = [mx.nd.empty(s, mx.gpu(0)) for s in shapes]
    for step in range(4):
        grads = [[rng.uniform(-0.4, 0.4, s).astype(np.float32) for _ in range(nsrc)] for s in shapes]
        kv.push(keys, [[mx.nd.array(a, mx.gpu(0)) for a in gs] for gs in grads])
        kv.pull(keys, out=outs)
        for k, s in enumerate(shapes):
            n = int(np.prod(s))
            deq = []
            for i in range(nsrc):
                comp = oracle.quantize_2bit(grads[k][i].ravel(), res[k][i], t)
                deq.append(oracle.dequantize_2bit(comp, n, t))
            want = oracle.reduce(deq, 'device')
            assert eq(outs[k].asnumpy().ravel(), want), (step, k)


def test_twobit_with_optimizer_and_errors(mx, oracle):
    rng = np.random.default_rng(60)
    n = 4099
    kv = mx.kv.create('device')
    kv.set_gradient_compression({'type': '2bit', 'threshold': 0.5})
    w = rng.uniform(-1, 1, n).astype(np.float32)
    kv.init(0, mx.nd.array(w, mx.gpu(0)))
    kv.set_optimizer(mx.optimizer.Test(rescale_grad=1.0))
    res = np.zeros(n, np.float32)
    out = mx.nd.empty((n,), mx.gpu(0))
    for step in range(3):
        g = rng.uniform(-1, 1, n).astype(np.float32)
        kv.pushpull(0, mx.nd.array(g, mx.gpu(0)), out=out)
        comp = oracle.quantize_2bit(g, res, 0.5)
        w = w + oracle.dequantize_2bit(comp, n, 0.5)          # Test optimizer: w += 1.0 * merged
        assert eq(out.asnumpy(), w), step
    with pytest.raises(mx.MXNetError, match="Unknown type"):
        kv.set_gradient_compression({'type': '1bit'})
    with pytest.raises(Exception, match="not supported"):
        mx.kv.create('local').set_gradient_compression({'type': '2bit'})
