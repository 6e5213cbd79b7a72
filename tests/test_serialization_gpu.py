"""NDArray::Save / Load binary format: the bytes python pickles of NDArrays -- hence
`Updater.get_states()` optimizer checkpoints and `kv.save_optimizer_states` files -- and
`mx.nd.save` files contain. Checked byte for byte against tests/golden/ndarray_raw_bytes.npz, which
oracle/gen_golden.py produced with the reference's own TShape::Save / Context::Save in
NDArray::Save's field order (src/ndarray/ndarray.cc:1596-1670), plus the legacy layouts Load must
still read (ndarray.cc:1672-1717)."""
import os
import pickle
import struct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ndarray_raw_bytes.npz")


def _raw(nd):
    return bytes(nd.__getstate__()['handle'])


def test_save_bytes_match_reference_format():
    import anand_mxnet_b200 as mx
    g = np.load(GOLD)
    a = mx.nd.array(g["dense_f32_gpu0_in"], mx.gpu(0))
    assert _raw(a) == g["dense_f32_gpu0_bytes"].tobytes()
    h = mx.nd.array(g["dense_f16_cpu_in"], mx.cpu(), np.float16)
    assert _raw(h) == g["dense_f16_cpu_bytes"].tobytes()
    i = mx.nd.array(g["dense_i64_cpu_in"], mx.cpu(), np.int64)
    assert _raw(i) == g["dense_i64_cpu_bytes"].tobytes()
    r = mx.nd.sparse.row_sparse_array((g["rsp_f32_gpu0_rows"], g["rsp_f32_gpu0_idx"]), shape=(6, 4),
                                      ctx=mx.gpu(0))
    assert _raw(r) == g["rsp_f32_gpu0_bytes"].tobytes()


def test_load_round_trip_and_legacy_layouts():
    import anand_mxnet_b200 as mx
    g = np.load(GOLD)
    for key, dtype in (("dense_f32_gpu0", np.float32), ("dense_f16_cpu", np.float16), ("dense_i64_cpu", np.int64)):
        nd = pickle.loads(pickle.dumps(mx.nd.array(g[key + "_in"], mx.gpu(0) if "gpu" in key else mx.cpu(), dtype)))
        assert nd.dtype == dtype and np.array_equal(nd.asnumpy(), g[key + "_in"])
        assert nd.context.device_type == ("gpu" if "gpu" in key else "cpu")
    r = pickle.loads(g["rsp_f32_gpu0_bytes"].tobytes() and pickle.dumps(
        mx.nd.sparse.row_sparse_array((g["rsp_f32_gpu0_rows"], g["rsp_f32_gpu0_idx"]), shape=(6, 4), ctx=mx.gpu(0))))
    assert r.stype == 'row_sparse' and np.array_equal(r.indices.asnumpy(), g["rsp_f32_gpu0_idx"])
    assert np.array_equal(r.data.asnumpy(), g["rsp_f32_gpu0_rows"])
    # V1 (magic 0xF993fac8, no storage type) and V0 (no magic: uint32 ndim + uint32 dims)
    vals = np.arange(6, dtype=np.float32)
    v1 = struct.pack('<I', 0xF993fac8) + struct.pack('<iqq', 2, 2, 3) + struct.pack('<iii', 1, 0, 0) + vals.tobytes()
    v0 = struct.pack('<III', 2, 2, 3) + struct.pack('<iii', 1, 0, 0) + vals.tobytes()
    for blob in (v1, v0):
        nd = mx.nd.NDArray.__new__(mx.nd.NDArray)
        nd.__setstate__({'handle': bytearray(blob)})
        assert nd.shape == (2, 3) and np.array_equal(nd.asnumpy().ravel(), vals)
    with pytest.raises(mx.MXNetError):
        nd = mx.nd.NDArray.__new__(mx.nd.NDArray)
        nd.__setstate__({'handle': bytearray(v1[:20])})          # truncated


def test_nd_save_load_file(tmp_path):
    import anand_mxnet_b200 as mx
    rng = np.random.default_rng(1)
    d = {'w': mx.nd.array(rng.uniform(-1, 1, (4, 3)).astype(np.float32), mx.gpu(0)),
         'b': mx.nd.array(rng.uniform(-1, 1, (3,)).astype(np.float32), mx.cpu())}
    f = str(tmp_path / "params.nd")
    mx.nd.save(f, d)
    back = mx.nd.load(f)
    assert sorted(back) == ['b', 'w']
    for k in d:
        assert np.array_equal(back[k].asnumpy(), d[k].asnumpy())
    raw = open(f, 'rb').read()
    assert raw[:16] == struct.pack('<QQ', 0x112, 0) and raw[16:24] == struct.pack('<Q', 2)
    mx.nd.save(f, [d['w']])
    assert isinstance(mx.nd.load(f), list)


def test_optimizer_state_checkpoint_round_trip(tmp_path):
    """kv.save_optimizer_states / load_optimizer_states (kvstore.py:552-582) on the callback route
    and on the fused route give the same continuation"""
    import anand_mxnet_b200 as mx
    rng = np.random.default_rng(2)
    shapes = [(5, 7), (33,)]
    w0 = [rng.uniform(-1, 1, s).astype(np.float32) for s in shapes]
    grads = [[rng.uniform(-1, 1, s).astype(np.float32) for s in shapes] for _ in range(4)]

    def run(resume_at, fname):
        kv = mx.kv.create('device')
        for k, w in enumerate(w0):
            kv.init(k, mx.nd.array(w, mx.gpu(0)))
        kv.set_optimizer(mx.optimizer.Adam(learning_rate=0.01, wd=0.01))
        outs = [mx.nd.zeros(s, mx.gpu(0)) for s in shapes]
        for t in range(4):
            if t == resume_at:
                kv.save_optimizer_states(fname, dump_optimizer=True)
                kv.load_optimizer_states(fname)
            kv.pushpull(list(range(len(shapes))), [mx.nd.array(g, mx.gpu(0)) for g in grads[t]], out=outs)
        return [o.asnumpy() for o in outs]
    f = str(tmp_path / "opt.states")
    a = run(None, f)
    b = run(2, f)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    states = pickle.loads(open(f, 'rb').read())
    assert isinstance(states, tuple) and len(states) == 2         # (states, optimizer)


def test_nccl_style_grouped_key_batches():
    """model._update_params_on_kvstore_nccl (python/mxnet/model.py:138-156): keys pushed and
    pulled in batches of 4 names with per-device value lists, store type 'nccl'"""
    import anand_mxnet_b200 as mx
    import torch
    ndev = min(torch.cuda.device_count(), 2)
    ctxs = [mx.gpu(i) for i in range(ndev)]
    rng = np.random.default_rng(3)
    names = ['p%d' % i for i in range(10)]
    shapes = [(3, 4), (17,), (2, 2, 2), (5,), (64, 3), (1,), (9, 9), (4,), (8, 2), (6,)]
    kv = mx.kv.create('nccl')
    assert kv.type == 'nccl'
    w0 = [rng.uniform(-1, 1, s).astype(np.float32) for s in shapes]
    for n, w in zip(names, w0):
        kv.init(n, mx.nd.array(w, ctxs[0]))
    grads = [[mx.nd.array(np.full(s, d + 1, np.float32), c) for d, c in enumerate(ctxs)] for s in shapes]
    args = [[mx.nd.zeros(s, c) for c in ctxs] for s in shapes]
    size, start = 4, 0
    while start < len(names):
        end = min(start + size, len(names))
        kv.push(names[start:end], grads[start:end], priority=-start)
        kv.pull(names[start:end], args[start:end], priority=-start)
        start = end
    total = sum(range(1, ndev + 1))
    for i, s in enumerate(shapes):
        for d in range(ndev):
            assert np.array_equal(args[i][d].asnumpy(), np.full(s, total, np.float32)), (names[i], d)
