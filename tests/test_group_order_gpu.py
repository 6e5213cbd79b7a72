"""Order of a key's values inside one grouped call (KVStoreLocal::GroupKVPairs, kvstore_local.h:377-407).

The reference sorts the call's (key, position) pairs with std::sort -- not stable -- so in a call of
more than 16 pairs (here 10 keys x 4 values) the values of a key reach the reduce in the order
libstdc++'s introsort leaves them (oracle: kvoracle.group_positions(..., 'reference'), pinned to the
reference's own function in tests/test_oracle.py). The product sums in call order by default;
B200KV_GROUP_ORDER=reference reproduces the reference's order. Both bit-exact against the oracle's
left fold over the respective order."""
import numpy as np
import pytest

import kvoracle as K

pytestmark = pytest.mark.gpu


def eq(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_group_order_call_and_reference(oracle, monkeypatch):
    import anand_mxnet_b200 as mx
    rng = np.random.default_rng(31)
    nkeys, nval, shape = 10, 4, (257,)
    keys = list(range(nkeys))
    grads = [[rng.uniform(-1, 1, shape).astype(np.float32) for _ in range(nval)] for _ in keys]
    flat_keys = np.repeat(np.arange(nkeys), nval)           # how the nested lists flatten: key-major
    flat_vals = [g for per_key in grads for g in per_key]
    results = {}
    for order in ('call', 'reference'):
        if order == 'reference':
            monkeypatch.setenv('B200KV_GROUP_ORDER', 'reference')
        else:
            monkeypatch.delenv('B200KV_GROUP_ORDER', raising=False)
        kv = mx.kv.create('device')
        for k in keys:
            kv.init(k, mx.nd.array(np.zeros(shape, np.float32), mx.gpu(0)))
        outs = [mx.nd.empty(shape, mx.gpu(0)) for _ in keys]
        kv.pushpull(keys, [[mx.nd.array(g, mx.gpu(0)) for g in per_key] for per_key in grads], out=outs)
        uniq, pos = K.group_positions(flat_keys, order)
        assert uniq == keys
        results[order] = [o.asnumpy() for o in outs]
        for k in keys:
            want = oracle.reduce([flat_vals[i] for i in pos[k]], 'device')
            assert eq(results[order][k], want.reshape(shape)), (order, k)
    # the two orders are different permutations for this call, and float sums notice
    assert K.group_positions(flat_keys, 'call') != K.group_positions(flat_keys, 'reference')
    assert any(not eq(a, b) for a, b in zip(results['call'], results['reference']))
